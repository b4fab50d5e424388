#!/bin/bash
# Multi-GPU evidence run of a round (inside gpurun --gpus N): S1 and S3 benches, N-rank == 1-rank check.
#   tools/evidence_multi.sh r2 8
tag=${1:-rX}; n=${2:-2}
out=gpurun_out
mkdir -p $out
run="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533"
$run bench.py --gpus $n --steps 10 --warmup 3 > $out/${tag}_bench_n$n.json 2> $out/${tag}_bench_n$n.err; echo "bench s1 n=$n: $?"
$run bench.py --gpus $n --steps 5 --warmup 3 --workload s3 > $out/${tag}_bench_s3_n$n.json 2> $out/${tag}_bench_s3_n$n.err; echo "bench s3 n=$n: $?"
$run tools/check_multirank_s3.py 64 > $out/${tag}_check_multirank_s3_n$n.log 2>&1; echo "check s3 n=$n: $?"
$run bench.py --impl reference --gpus $n --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm_n$n.json 2> $out/${tag}_bench_ref_n$n.err; echo "ref n=$n: $?"
tail -2 $out/${tag}_check_multirank_s3_n$n.log
