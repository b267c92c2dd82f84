#!/bin/bash
# round-2 session a: GPU tests, api workload (1 GPU), default bench
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -5 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --workload api --steps 3 > gpurun_out/r2a_api_n1.json 2> gpurun_out/r2a_api_n1.err; echo "api rc=$?"
tail -c 1500 gpurun_out/r2a_api_n1.json; tail -5 gpurun_out/r2a_api_n1.err
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2a_bench_n1.json
