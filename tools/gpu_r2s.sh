#!/bin/bash
# round-2 session s: default bench line with the steady-state coreset record and 5-call API medians
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2s_bench_n1.json 2> gpurun_out/r2s_bench_n1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2s_bench_n1.err
