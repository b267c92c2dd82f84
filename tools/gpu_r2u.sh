#!/bin/bash
# round-2 session u: per-line executed-instruction counts of the MPE stream kernel and of the fused MPE launch
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_stream_kernel -s 3 -c 1 -f -o gpurun_out/r2u_stream_mpe \
  python bench.py --workload scores --scores-only "PeaksOp<0>" > gpurun_out/r2u_ncu_stream.log 2>&1; echo "ncu stream rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_pool_fused_kernel -s 3 -c 1 -f -o gpurun_out/r2u_fused_mpe \
  python bench.py --workload scores --scores-only "MPE> in ONE launch" > gpurun_out/r2u_ncu_fused.log 2>&1; echo "ncu fused rc=$?"
