#!/bin/bash
# round-2 session v2: default line (no extra records) on N = $1 GPUs, final tree
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-extra --cpu-frames 0 > gpurun_out/r2v2_bench_n$N.json 2> gpurun_out/r2v2_bench_n$N.err; echo "bench rc=$?"
tail -c 200 gpurun_out/r2v2_bench_n$N.json
