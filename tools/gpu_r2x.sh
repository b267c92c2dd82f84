#!/bin/bash
# round-2 session x: scan with the OR on the FMA pipe; all score kernels
mkdir -p gpurun_out
T=${1:-x}
timeout 600 python bench.py --workload scores > gpurun_out/r2${T}_scores.json 2> gpurun_out/r2${T}_scores.err; echo "scores rc=$?"
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fused or scored or score or peak or mpe or bsb" > gpurun_out/r2${T}_pytest.log 2>&1; tail -2 gpurun_out/r2${T}_pytest.log
