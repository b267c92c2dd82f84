#!/bin/bash
# round-2 session o2: compute-sanitizer racecheck + memcheck over the map-stream tests (dynamic 14-stage ring, in-place BSB rewrite, XOR row blocks)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peak or plateau" > gpurun_out/r2o2_racecheck.log 2>&1; echo "racecheck rc=$?"
tail -4 gpurun_out/r2o2_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "peak or plateau or scored or softargmax or xe or hp" > gpurun_out/r2o2_memcheck.log 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r2o2_memcheck.log
