# Round-1f last GPU session: the tests added after the fourth call, a launch list of the default bench command, the
# one-GPU hybrid line.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests -q -m gpu -k "softargmax or strategy or sal or kmeans or golden" > gpurun_out/t_last.log 2>&1
echo "rc=$?" >> gpurun_out/t_last.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_default.csv \
  python bench.py --steps 2 --warmup 3 --cpu-frames 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 200 python bench.py --workload hybrid > gpurun_out/bench_hybrid_n1.json 2> gpurun_out/bench_hybrid_n1.err
tail -n 3 gpurun_out/t_last.log gpurun_out/bench_hybrid_n1.err
