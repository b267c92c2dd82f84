set -x
B="python bench.py --workload scores --resident-frames 4096"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:map_stream --launch-skip 25 --launch-count 3 -o gpurun_out/prof_stream -f $B > gpurun_out/prof_stream.log 2>&1
tail -3 gpurun_out/prof_stream.log
