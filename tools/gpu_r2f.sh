#!/bin/bash
# round-2 session f: launch list of one 8-GPU shard of C4 (125k x 2048) and of the d = 57 pool, scored kernels after the code-size cuts
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2f_launches_coreset_shard.csv \
  python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2f_ncu_coreset.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2f_coreset_shard.json 2> gpurun_out/r2f_coreset_shard.err; echo "shard rc=$?"
tail -c 600 gpurun_out/r2f_coreset_shard.json
timeout 600 python bench.py --workload scores --scores-only fused > gpurun_out/r2f_scores.json 2> gpurun_out/r2f_scores.err; echo "scores rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peak or fused or stream or mpe or plateau" > gpurun_out/r2f_pytest.log 2>&1; tail -3 gpurun_out/r2f_pytest.log
