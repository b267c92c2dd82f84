#!/bin/bash
# round-2 session e2: XOR-rotated row blocks (arg-max sweep, HP, BSB softmax pass); all GPU tests + score kernels
mkdir -p gpurun_out
T=${1:-e2}
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2${T}_pytest.log; tail -3 gpurun_out/r2${T}_pytest.log
timeout 600 python bench.py --workload scores > gpurun_out/r2${T}_scores.json 2> gpurun_out/r2${T}_scores.err; echo "scores rc=$?"
