#!/bin/bash
# round-2 session w: warp split of the scored fused launch (decode / RANSAC warps)
mkdir -p gpurun_out
T=${1:-w}
timeout 600 python bench.py --workload scores --scores-only "score_pool" > gpurun_out/r2${T}_scores.json 2> gpurun_out/r2${T}_scores.err; echo "scores rc=$?"
for s in 1 2; do
MVAL_FUSED_SHAPE=$s MVAL_SCORED_SPLIT=0 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or scored or score" > gpurun_out/r2${T}_pytest_shape$s.log 2>&1; tail -2 gpurun_out/r2${T}_pytest_shape$s.log
done
