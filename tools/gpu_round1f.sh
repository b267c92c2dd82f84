# Round-1f GPU session (one B200): parity of the scored fused pass, per-kernel roofline lines, ncu captures, full GPU
# test suite, default bench line.  Every step has its own timeout and writes to gpurun_out/; later steps run even if an
# earlier one fails.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/smi.txt 2>&1
MVAL_DEBUG_SYNC=1 timeout 400 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "fused or map_stream or empty_inputs" > gpurun_out/t_fused.log 2>&1
echo "rc=$?" >> gpurun_out/t_fused.log
timeout 300 python -m pytest tests/test_gpu_strategy.py -q -m gpu > gpurun_out/t_strategy.log 2>&1
echo "rc=$?" >> gpurun_out/t_strategy.log
timeout 300 python bench.py --workload scores > gpurun_out/bench_scores.json 2> gpurun_out/bench_scores.err
B="python bench.py --workload scores --resident-frames 4096"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_pool_fused --launch-skip 4 --launch-count 1 \
  -o gpurun_out/prof_fused_HP -f $B --scores-only "fused_kernel<HP> (" > gpurun_out/prof_fused_HP.log 2>&1
timeout 500 python -m pytest tests -q -m gpu > gpurun_out/t_all.log 2>&1
echo "rc=$?" >> gpurun_out/t_all.log
timeout 400 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"score_pool_fused|map_stream|decode_argmax" -c 240 --csv \
  --log-file gpurun_out/launches_scores.csv $B > /dev/null 2>&1
for kind in MPE BSB; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_pool_fused --launch-skip 4 --launch-count 1 \
    -o gpurun_out/prof_fused_$kind -f $B --scores-only "fused_kernel<$kind> (" > gpurun_out/prof_fused_$kind.log 2>&1
done
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ls -la gpurun_out
tail -3 gpurun_out/t_fused.log gpurun_out/t_strategy.log gpurun_out/t_all.log
