#!/bin/bash
# round-2 session z: per-line executed-instruction counts of the BSB stream kernel and of the arg-max + MPE stream kernel
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_stream_kernel -s 3 -c 1 -f -o gpurun_out/r2z_stream_bsb \
  python bench.py --workload scores --scores-only "PeaksOp<1>" > gpurun_out/r2z_ncu_bsb.log 2>&1; echo "ncu bsb rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:map_stream_kernel -s 3 -c 1 -f -o gpurun_out/r2z_stream_argmax_mpe \
  python bench.py --workload scores --scores-only "argmax+MPE" > gpurun_out/r2z_ncu_ampe.log 2>&1; echo "ncu argmax+mpe rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ransac_vote_kernel -s 3 -c 1 -f -o gpurun_out/r2z_vote \
  python bench.py --workload scores --scores-only "argmax+MPE" > gpurun_out/r2z_ncu_vote.log 2>&1; echo "ncu vote rc=$?"
