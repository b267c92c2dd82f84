#!/bin/bash
# round-2 session c: GPU tests (flow goldens, plateaus, threshold votes, pipeline flags, segments), rank inversions,
# C4 with the staged replay, ncu --set full of the fused kernel (segments) for roofline.traffic
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -6 gpurun_out/r2c_pytest.log
timeout 900 python tools/rank_inversions.py 100000 > gpurun_out/r2c_rank_inversions.json 2> gpurun_out/r2c_rank_inversions.err; echo "rank rc=$?"
tail -c 1200 gpurun_out/r2c_rank_inversions.json; tail -3 gpurun_out/r2c_rank_inversions.err
timeout 600 python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2c_coreset_c4_n1.json 2> gpurun_out/r2c_coreset_c4_n1.err; echo "c4 rc=$?"
tail -c 700 gpurun_out/r2c_coreset_c4_n1.json
timeout 600 python bench.py --workload coreset --coreset-rows 1000000 --coreset-dim 57 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2c_coreset_d57_n1.json 2> gpurun_out/r2c_coreset_d57_n1.err; echo "d57 rc=$?"
tail -c 500 gpurun_out/r2c_coreset_d57_n1.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_pool_fused -c 2 -o gpurun_out/r2c_fused_segments \
  python bench.py --steps 1 --warmup 3 --no-extra --e2e-steps 1 --cpu-frames 0 --pool-frames 16384 > gpurun_out/r2c_ncu_full.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep 2>/dev/null
