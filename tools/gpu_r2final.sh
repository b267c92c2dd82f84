#!/bin/bash
# round-2 final verification of the round-2 tree: GPU tests, API strategies, score kernels, default bench, reference arm, smoke
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2final_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2final_pytest.log
tail -3 gpurun_out/r2final_pytest.log
timeout 600 python bench.py --workload api --steps 3 --api-variants HP/AL,MPE/AL,BSB/AL,HP/SAL > gpurun_out/r2final_api_scores_n1.json 2> gpurun_out/r2final_api_scores_n1.err; echo "api rc=$?"
timeout 600 python bench.py --workload scores > gpurun_out/r2final_scores.json 2> gpurun_out/r2final_scores.err; echo "scores rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2final_bench_n1.json 2> gpurun_out/r2final_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_reference_arm.json 2> gpurun_out/r2final_reference_arm.err; echo "ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2final_smoke.log 2>&1; tail -1 gpurun_out/r2final_smoke.log
