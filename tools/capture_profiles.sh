# Round-2 evidence, run under gpurun on one B200 (each plain command first; numbers under ncu are never bench values):
#   GPU parity suite, bench (product arm + reference arm), ncu launch list + one --set full capture per kernel
#   (guided table kernel on the FULL config-3 launch, literal scan, FP32 mode, small-G mode), config report, fuzz.
set -x
O=gpurun_out
python -m pytest tests -m gpu -q > $O/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_r2.log
python bench.py > $O/bench_r2_n1.json 2> $O/bench_r2_n1.err || exit 1
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_r2_reference_arm.json 2> $O/bench_r2_reference_arm.err
B="python bench.py --steps 2 --warmup 3 --no-cpu --no-extras"
$B > $O/plain_guided.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_r2_guided.csv $B > $O/ncu_l_guided.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tofrt_table_kernel -s 3 -c 1 -o $O/prof_r2_guided -f $B > $O/ncu_f_guided.log 2>&1
$B --search scan > $O/plain_scan.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tofrt_table_kernel -s 3 -c 1 -o $O/prof_r2_scan -f $B --search scan > $O/ncu_f_scan.log 2>&1
F="python bench.py --steps 1 --warmup 3 --no-cpu --no-3d"
ncu --set full --clock-control none --import-source on -k regex:tofrt_table_kernel_f32 -s 1 -c 1 -o $O/prof_r2_f32 -f $F > $O/ncu_f_f32.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tofrt_table_kernel -s 3 -c 1 -o $O/prof_r2_smallg -f python tools/_variants/cfg1.py > $O/ncu_f_smallg.log 2>&1
python tools/ncu_summary.py $O/ncu_r2_table_kernel.json guided=$O/prof_r2_guided.ncu-rep scan=$O/prof_r2_scan.ncu-rep f32=$O/prof_r2_f32.ncu-rep small_g=$O/prof_r2_smallg.ncu-rep > $O/ncu_summary.log 2>&1
python tools/config_report.py --big > $O/configs_r2.json 2> $O/configs_r2.err
FUZZ_ORACLE_GEOMS=${FUZZ_ORACLE_GEOMS:-400000} FUZZ_MODE_GEOMS=${FUZZ_MODE_GEOMS:-20000000} FUZZ_SETUPS=${FUZZ_SETUPS:-30} python tools/fuzz_parity.py > $O/fuzz_r2.json 2> $O/fuzz_r2.err
tail -2 $O/pytest_gpu_r2.log; head -c 300 $O/bench_r2_n1.json; tail -4 $O/ncu_summary.log; tail -30 $O/fuzz_r2.json
