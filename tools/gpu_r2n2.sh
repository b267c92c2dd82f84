#!/bin/bash
# round-2 session n2: BSB's softmax pass reuses the arg-max sweep's row maxima
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strategy.py -m gpu -x -q > gpurun_out/r2n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n2_pytest.log; tail -3 gpurun_out/r2n2_pytest.log
timeout 600 python bench.py --workload scores --scores-only "BSB" > gpurun_out/r2n2_scores.json 2> gpurun_out/r2n2_scores.err; echo "scores rc=$?"
