#!/bin/bash
# round-2 session m2: per-line profile of the final BSB stream kernel (arg-max + BSB, the scored default)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_stream_kernel -s 3 -c 1 -f -o gpurun_out/r2m2_stream_argmax_bsb \
  python bench.py --workload scores --scores-only "argmax+BSB" > gpurun_out/r2m2_ncu.log 2>&1; echo "ncu rc=$?"
