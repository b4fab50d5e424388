#!/bin/bash
# kernel time of the frame kernel for one or more builds (inside gpurun)
for so in "${@:-contact_map_b200/libcmb200.so}"; do
  echo "== $so"; CMB200_SO=$so python tools/profile_s1.py 10000 4 | tail -3 | head -2
done
