#!/bin/bash
# round-2 session i2: _compute_sal_dict with a HOST loader (pinned batches, upload of batch k+1 on a copy stream under batch k) vs resident batches
mkdir -p gpurun_out
timeout 600 python tools/bench_sal_dict.py 8192 512 TRIANGULATION pinned > gpurun_out/r2i2_sal_dict_pinned.log 2>&1; tail -3 gpurun_out/r2i2_sal_dict_pinned.log
timeout 600 python tools/bench_sal_dict.py 8192 2048 TRIANGULATION pinned > gpurun_out/r2i2_sal_dict_pinned_b2048.log 2>&1; tail -3 gpurun_out/r2i2_sal_dict_pinned_b2048.log
