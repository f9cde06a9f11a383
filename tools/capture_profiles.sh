# ncu evidence (run under gpurun): launch list + one --set full capture per kernel; each plain command first.
#   CAPTURE_ONLY="guided" limits the table-kernel captures to one search mode and skips the 3-D kernel
set -x
B="python bench.py --steps 2 --warmup 3 --grid 128 --no-cpu"
for m in ${CAPTURE_ONLY:-guided scan}; do
  $B --search $m > gpurun_out/plain_$m.log 2>&1 || exit 1
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1_$m.csv $B --search $m > gpurun_out/ncu_l_$m.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:tofrt_table_kernel -s 3 -c 1 -o gpurun_out/prof_r1_${m}_v3 -f $B --search $m > gpurun_out/ncu_f_$m.log 2>&1
done
if [ -z "$CAPTURE_ONLY" ]; then
C="python tools/config5_report.py --slabs 4 --steps 1"
$C > gpurun_out/c5_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1_config5.csv $C > gpurun_out/ncu_c5_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tofrt_fermat3d_kernel -c 1 -s 1 -o gpurun_out/prof_r1_fermat3d -f $C > gpurun_out/ncu_c5_f.log 2>&1
fi
