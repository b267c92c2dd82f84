#!/bin/bash
# round-2 session j2: ranking agreement of HP / MPE / BSB with the oracle on 100 000 frames, final kernels
mkdir -p gpurun_out
timeout 1200 python tools/rank_inversions.py 100000 > gpurun_out/r2j2_rank_inversions.json 2> gpurun_out/r2j2_rank_inversions.err; echo "rank rc=$?"
tail -c 600 gpurun_out/r2j2_rank_inversions.json; tail -3 gpurun_out/r2j2_rank_inversions.err
