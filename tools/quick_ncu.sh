#!/bin/bash
# instruction count / issue utilisation / duration of the frame kernel for one or more builds
# usage: tools/quick_ncu.sh [so ...]   (inside gpurun)
for so in "${@:-contact_map_b200/libcmb200.so}"; do
  echo "== $so"
  CMB200_SO=$so python tools/profile_s1.py 2000 3 | tail -2
  CMB200_SO=$so ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio,sm__warps_active.avg.pct_of_peak_sustained_active \
     --clock-control none -k regex:k_frames_small -c 1 -s 2 python tools/profile_s1.py 2000 3 2>&1 | grep -E "gpu__time|inst_executed|issue_active|no_instruction|warps_active"
done
