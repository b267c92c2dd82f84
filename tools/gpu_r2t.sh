#!/bin/bash
# round-2 session t: launch list of the default bench command as it is now (one fused launch per step + the extra records);
# coreset record with the round's own centres
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r2t_launches_default_bench.csv \
  python bench.py --steps 2 --warmup 1 > gpurun_out/r2t_ncu_default.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2t_coreset_c4_n1.json 2> gpurun_out/r2t_coreset_c4_n1.err; echo "c4 rc=$?"
timeout 600 python bench.py --workload coreset --coreset-rows 125000 --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2t_coreset_shard.json 2> gpurun_out/r2t_coreset_shard.err; echo "shard rc=$?"
