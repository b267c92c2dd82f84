#!/bin/bash
# round-2 session c2: vote kernel at 3 blocks / SM; rigs sized to multiples of the SM count; default bench line
mkdir -p gpurun_out
T=${1:-c2}
timeout 600 python bench.py --workload scores --scores-only "RANSAC launches" > gpurun_out/r2${T}_scores_split.json 2> gpurun_out/r2${T}_scores_split.err; echo "scores rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2${T}_bench_n1.json 2> gpurun_out/r2${T}_bench_n1.err; echo "bench rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ransac or triang or golden or near" > gpurun_out/r2${T}_pytest.log 2>&1; tail -2 gpurun_out/r2${T}_pytest.log
