#!/bin/bash
# round-2 session h: tests after the replay / recheck rewrites, coreset shard + C4 timings, launch list of the shard
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
tail -4 gpurun_out/r2h_pytest.log
timeout 600 python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2h_coreset_shard.json 2> gpurun_out/r2h_coreset_shard.err; echo "shard rc=$?"
timeout 600 python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2h_coreset_c4_n1.json 2> gpurun_out/r2h_coreset_c4_n1.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload coreset --coreset-rows 1000000 --coreset-dim 57 --coreset-pad 64 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2h_coreset_d57.json 2> gpurun_out/r2h_coreset_d57.err; echo "d57 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2h_launches_coreset_shard.csv \
  python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2h_ncu_coreset.log 2>&1; echo "ncu rc=$?"
