#!/bin/bash
# round-2 session t2: launch list of the default bench command, final tree
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2t2_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/r2t2_ncu_default.log 2>&1; echo "ncu rc=$?"
