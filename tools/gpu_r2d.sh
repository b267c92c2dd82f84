#!/bin/bash
# round-2 session d: tests, default bench (sleep back-off), scored kernels, coreset at the reference's feature sizes with / without padding
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -4 gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench_n1.json 2> gpurun_out/r2d_bench_n1.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2d_bench_n1.json; tail -3 gpurun_out/r2d_bench_n1.err
timeout 600 python bench.py --workload scores --scores-only fused > gpurun_out/r2d_scores.json 2> gpurun_out/r2d_scores.err; echo "scores rc=$?"
for cfg in "57 0" "57 64" "126 0" "126 128"; do
  set -- $cfg
  timeout 600 python bench.py --workload coreset --coreset-rows 1000000 --coreset-dim $1 --coreset-pad $2 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 \
    > gpurun_out/r2d_coreset_d$1_pad$2.json 2> gpurun_out/r2d_coreset_d$1_pad$2.err; echo "coreset d=$1 pad=$2 rc=$?"
done
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_reference_arm.json 2> gpurun_out/r2d_reference_arm.err; echo "ref rc=$?"
