# Round-1f two-GPU sanity session: the 2-rank NCCL tests and the default / hybrid bench lines under torchrun.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_multirank.py -q -m gpu > gpurun_out/t_multirank.log 2>&1
echo "rc=$?" >> gpurun_out/t_multirank.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --workload hybrid --verify > gpurun_out/bench_hybrid_n2.json 2> gpurun_out/bench_hybrid_n2.err
timeout 200 $TR --master-port 29513 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err
tail -n 3 gpurun_out/t_multirank.log gpurun_out/bench_n2.err gpurun_out/bench_hybrid_n2.err gpurun_out/bench_ref_n2.err
