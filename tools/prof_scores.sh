set -x
B="python bench.py --workload scores --resident-frames 4096"
for k in decode_softargmax score_peaks_kernel score_xe score_hp; do
  skip=1; cnt=1
  if [ $k = score_peaks_kernel ]; then cnt=10; fi
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip --launch-count $cnt -o gpurun_out/prof_$k -f $B > gpurun_out/prof_$k.log 2>&1
done
ls -la gpurun_out
