# ncu --set full capture of one launch of each persistent map-stream kernel (bench.py --workload scores calls each op
# 3 + 1 + 5 times: soft-arg-max, HP, MPE, BSB, XE in that order; 8 launches apart... the XE op also launches xe_frame_reduce)
set -x
B="python bench.py --workload scores --resident-frames 4096"
for skip in 4 13 22 31 40; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:map_stream --launch-skip $skip --launch-count 1 -o gpurun_out/prof_stream_$skip -f $B > gpurun_out/prof_stream_$skip.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_scores.csv $B > /dev/null 2>&1
ls -la gpurun_out
