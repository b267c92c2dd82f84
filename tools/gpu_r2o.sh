#!/bin/bash
# round-2 session o: scored split path (tests + scores workload), default bench, api with HP / MPE / BSB
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2o_pytest.log
tail -3 gpurun_out/r2o_pytest.log
timeout 600 python bench.py --workload scores --scores-only "MPE" > gpurun_out/r2o_scores_mpe.json 2> gpurun_out/r2o_scores.err; echo "scores rc=$?"
timeout 600 python bench.py --workload scores --scores-only "BSB" > gpurun_out/r2o_scores_bsb.json 2>> gpurun_out/r2o_scores.err; echo "scores rc=$?"
timeout 600 python bench.py --workload scores --scores-only "HP" > gpurun_out/r2o_scores_hp.json 2>> gpurun_out/r2o_scores.err; echo "scores rc=$?"
timeout 600 python bench.py --workload api --steps 3 --api-variants HP/AL,MPE/AL,BSB/AL > gpurun_out/r2o_api_scores_n1.json 2> gpurun_out/r2o_api_scores_n1.err; echo "api rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench_n1.json 2> gpurun_out/r2o_bench_n1.err; echo "bench rc=$?"
