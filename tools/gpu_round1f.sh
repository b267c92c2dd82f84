# Round-1f GPU session (one B200): parity of the scored fused pass, per-kernel roofline lines, the strategy-level effect,
# full GPU test suite, default bench line.  Every step has its own timeout and writes to gpurun_out/; later steps run
# even if an earlier one fails.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MVAL_DEBUG_SYNC=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused or map_stream or empty_inputs or hp or mpe or decode" > gpurun_out/t_fused.log 2>&1
echo "rc=$?" >> gpurun_out/t_fused.log
timeout 300 python bench.py --workload scores > gpurun_out/bench_scores.json 2> gpurun_out/bench_scores.err
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
echo "rc=$?" >> gpurun_out/t_all.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
for s in TRIANGULATION HP MPE BSB; do
  timeout 120 python tools/bench_sal_dict.py 4096 64 $s resident >> gpurun_out/sal_dict.log 2>&1
done
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
B="python bench.py --workload scores --resident-frames 4096"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_pool_fused|map_stream|decode_argmax" -c 240 --csv \
  --log-file gpurun_out/launches_scores.csv $B > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_pool_fused --launch-skip 4 --launch-count 1 \
    -o gpurun_out/prof_fused_plain -f $B --scores-only "fused_kernel (a1" > gpurun_out/prof_fused_plain.log 2>&1
ls -la gpurun_out
tail -n 3 gpurun_out/t_fused.log gpurun_out/t_all.log gpurun_out/sal_dict.log gpurun_out/smoke.log
