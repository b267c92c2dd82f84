#!/bin/bash
# round-2 session b: new GPU tests, default bench with extras (N = 1), launch list of the default bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -4 gpurun_out/r2b_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2b_bench_n1.json 2> gpurun_out/r2b_bench_n1.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2b_bench_n1.json; tail -5 gpurun_out/r2b_bench_n1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 3 --no-extra --e2e-steps 1 --cpu-frames 0 > gpurun_out/r2b_ncu_bench.log 2>&1; echo "ncu rc=$?"
