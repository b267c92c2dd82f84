#!/bin/bash
# round-2 session n: coreset kernels after the recheck rewrite; source-level capture of the replay CTA
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strategy.py -m gpu -x -q -k "kcenter or coreset or selection" > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
timeout 600 python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2n_coreset_shard.json 2> gpurun_out/r2n_coreset_shard.err; echo "shard rc=$?"
timeout 600 python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2n_coreset_c4_n1.json 2> gpurun_out/r2n_coreset_c4_n1.err; echo "c4 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2n_launches_coreset_shard.csv \
  python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2n_ncu_coreset.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload coreset --coreset-rows 1000000 --coreset-dim 57 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 --coreset-pad 64 > gpurun_out/r2n_coreset_d57.json 2> gpurun_out/r2n_coreset_d57.err; echo "d57 rc=$?"
timeout 600 python bench.py --workload coreset --coreset-data clustered --coreset-rows 250000 --coreset-labeled 1000 --coreset-budget 4000 --cpu-frames 0 > gpurun_out/r2n_coreset_clustered.json 2> gpurun_out/r2n_coreset_clustered.err; echo "clustered rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest_all.log 2>&1; tail -3 gpurun_out/r2n_pytest_all.log
