#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --workload scores --scores-only "MPE" > gpurun_out/r2d2_scores_mpe.json 2> gpurun_out/r2d2_scores_mpe.err; echo "scores rc=$?"
timeout 600 python bench.py --workload scores --scores-only "MPE" > gpurun_out/r2d2_scores_mpe_b.json 2> gpurun_out/r2d2_scores_mpe_b.err; echo "scores rc=$?"
