#!/bin/bash
# One-GPU evidence run of a round (inside gpurun): tests, benches, launch lists, full ncu captures.
#   tools/evidence_n1.sh r2      -> gpurun_out/r2_*
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; echo "pytest exit $?" >> $out/${tag}_pytest_gpu.log
tail -2 $out/${tag}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $out/${tag}_smoke.log 2>&1; echo "smoke exit $?"
python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2> $out/${tag}_bench_ref.err; echo "ref $?"
python bench.py > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err; echo "bench $?"
python bench.py --workload s3 --steps 5 > $out/${tag}_bench_s3_n1.json 2> $out/${tag}_bench_s3_n1.err; echo "bench s3 $?"
# launch lists (per-launch durations, serialised, cold cache: shares only)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-legs --no-cpu-baseline --no-e2e > $out/ncu_bench.log 2>&1; echo "launches $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $out/${tag}_launches_s3.csv \
    python tools/profile_s1.py 512 2 s3 > $out/ncu_s3l.log 2>&1; echo "launches s3 $?"
# full captures of the dominant kernel
ncu --set full --clock-control none --import-source on -k regex:vl_frames -c 1 -s 2 -f -o $out/${tag}_vl_frames_s1 \
    python tools/profile_s1.py 10000 3 > $out/ncu_s1.log 2>&1; echo "ncu s1 $?"
ncu --set full --clock-control none --import-source on -k regex:vl_frames -c 1 -s 2 -f -o $out/${tag}_vl_frames_s3 \
    python tools/profile_s1.py 512 3 s3 > $out/ncu_s3.log 2>&1; echo "ncu s3 $?"
