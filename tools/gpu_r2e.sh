#!/bin/bash
# round-2 multi-GPU session: N = $1 ranks (torchrun): NCCL tests, default bench line (with extras), api workload, C4 coreset
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2e_topo_n$N.txt 2>&1
if [ "$N" = "2" ]; then
  timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2e_pytest_multirank.log 2>&1; echo "multirank rc=$?" >> gpurun_out/r2e_pytest_multirank.log
  tail -4 gpurun_out/r2e_pytest_multirank.log
fi
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2e_bench_n$N.json 2> gpurun_out/r2e_bench_n$N.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2e_bench_n$N.json; tail -3 gpurun_out/r2e_bench_n$N.err
timeout 600 $TR bench.py --gpus $N --workload api --steps 3 > gpurun_out/r2e_api_n$N.json 2> gpurun_out/r2e_api_n$N.err; echo "api rc=$?"
tail -c 300 gpurun_out/r2e_api_n$N.json; tail -3 gpurun_out/r2e_api_n$N.err
timeout 600 $TR bench.py --gpus $N --workload hybrid --steps 3 --verify > gpurun_out/r2e_hybrid_n$N.json 2> gpurun_out/r2e_hybrid_n$N.err; echo "hybrid rc=$?"
tail -c 400 gpurun_out/r2e_hybrid_n$N.json; tail -3 gpurun_out/r2e_hybrid_n$N.err
