#!/bin/bash
# round-2 session q: instruction-footprint cuts in the scored kernels (out-of-line Jacobi sweeps, scan rolled 6 x 5 + 4)
mkdir -p gpurun_out
T=${1:-q}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_strategy.py -m gpu -x -q > gpurun_out/r2${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${T}_pytest.log
tail -3 gpurun_out/r2${T}_pytest.log
timeout 600 python bench.py --workload scores > gpurun_out/r2${T}_scores.json 2> gpurun_out/r2${T}_scores.err; echo "scores rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 --no-extra > gpurun_out/r2${T}_bench_n1.json 2> gpurun_out/r2${T}_bench_n1.err; echo "bench rc=$?"
cat gpurun_out/r2${T}_scores.json | cut -c1-1500
