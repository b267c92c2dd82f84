#!/bin/bash
# round-2 final check on 2 GPUs: NCCL tests + default bench line under torchrun
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multirank.py -x -q > gpurun_out/r2l2_pytest_multirank.log 2>&1; echo "multirank rc=$?" >> gpurun_out/r2l2_pytest_multirank.log
tail -3 gpurun_out/r2l2_pytest_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2l2_bench_n2.json 2> gpurun_out/r2l2_bench_n2.err; echo "bench rc=$?"
tail -c 200 gpurun_out/r2l2_bench_n2.json
