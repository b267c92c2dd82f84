#!/bin/bash
# round-2 final multi-GPU check: N = $1 ranks (torchrun): default bench line with its extra records, the API workload for all strategies
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2f2_bench_n$N.json 2> gpurun_out/r2f2_bench_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2f2_bench_n$N.json; tail -3 gpurun_out/r2f2_bench_n$N.err
timeout 600 $TR bench.py --gpus $N --workload api --steps 3 --api-variants TRIANGULATION/AL,CORESET/AL,TRIANGULATION/SAL,HP/AL,MPE/AL,BSB/AL > gpurun_out/r2f2_api_n$N.json 2> gpurun_out/r2f2_api_n$N.err; echo "api rc=$?"
tail -c 300 gpurun_out/r2f2_api_n$N.json; tail -3 gpurun_out/r2f2_api_n$N.err
