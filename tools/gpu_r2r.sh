#!/bin/bash
# round-2 session r: what C4 on ONE GPU (1 M x 2048) is made of; full capture of one screen launch at that size
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2r_launches_coreset_c4_n1.csv \
  python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2r_ncu_coreset.log 2>&1; echo "ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:kc_screen_tc2 -s 12 -c 1 -f -o gpurun_out/r2r_screen_c4 \
  python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2r_ncu_screen.log 2>&1; echo "ncu full rc=$?"
timeout 600 python bench.py --workload coreset --coreset-labeled 1000 --coreset-budget 10000 --cpu-frames 0 > gpurun_out/r2r_coreset_c4_n1.json 2> gpurun_out/r2r_coreset_c4_n1.err; echo "c4 rc=$?"
