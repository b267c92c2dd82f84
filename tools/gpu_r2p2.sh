#!/bin/bash
# round-2 session p2: sustained default step with either arg-max flavour of the unscored fused pass (A/B/A/B on one box)
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --cpu-frames 0 > gpurun_out/r2p2_scan_$i.json 2> gpurun_out/r2p2_scan_$i.err; echo "scan $i rc=$?"
MVAL_ROW_ARGMAX=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-extra --cpu-frames 0 > gpurun_out/r2p2_row_$i.json 2> gpurun_out/r2p2_row_$i.err; echo "row $i rc=$?"
done
