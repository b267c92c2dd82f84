#!/bin/bash
# round-2 session b2: compute-sanitizer memcheck over the map-score / fused parity tests (new scan + walk code)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peak or mpe or bsb or scored or plateau" > gpurun_out/r2b2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -5 gpurun_out/r2b2_memcheck.log
grep -c "Invalid\|ERROR SUMMARY" gpurun_out/r2b2_memcheck.log
