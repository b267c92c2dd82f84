#!/bin/bash
# round-2 session q2: arg-max flavour of the unscored fused pass on SHORT launches (API batches of 8 192 frames, 16 384-frame launches), A/B/A/B on one box
mkdir -p gpurun_out
for i in 1 2; do
for f in 0 1; do
MVAL_ROW_ARGMAX=$f timeout 600 python bench.py --workload api --steps 5 --api-variants TRIANGULATION/AL > gpurun_out/r2q2_api_f${f}_$i.json 2> gpurun_out/r2q2_api_f${f}_$i.err; echo "api f$f $i rc=$?"
done
done
timeout 600 python bench.py --workload scores --scores-only "score_pool_fused_kernel" > gpurun_out/r2q2_scores.json 2> gpurun_out/r2q2_scores.err
