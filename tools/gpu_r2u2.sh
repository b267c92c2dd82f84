#!/bin/bash
# round-2 session u2: full capture of the unscored fused kernel in the flavour the default 100 000-frame step now uses (lane = row arg-max), for roofline.traffic
mkdir -p gpurun_out
MVAL_ROW_ARGMAX=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_pool_fused -c 2 -f -o gpurun_out/r2u2_fused_row \
  python bench.py --steps 1 --warmup 3 --no-extra --e2e-steps 1 --cpu-frames 0 --pool-frames 16384 > gpurun_out/r2u2_ncu_full.log 2>&1; echo "ncu rc=$?"
