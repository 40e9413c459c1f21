# final evidence of round 2: the default bench line, then ncu launch list + full captures (fp64 and fp32 bulk kernels) of a
# short run of the same workload (0.02 s per string so that ncu's replays stay short)
set -x
python bench.py > gpurun_out/r02_final_bench.log 2> gpurun_out/r02_final_bench.err; echo "bench rc=$?"
B="python bench.py --steps 1 --warmup 1 --fp32-steps 1 --length 0.02 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped"
SFDTD_VERBOSE=1 $B > gpurun_out/r02f_prof_plain.log 2> gpurun_out/r02f_prof_plain.err || exit 1
grep -h "bucket" gpurun_out/r02f_prof_plain.err | head -24
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02f.csv $B > gpurun_out/r02f_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"step_kernel<double, 16, 4" --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r02f_f64 $B > gpurun_out/r02f_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"step_kernel<float, 16, 4" --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r02f_f32 $B > gpurun_out/r02f_ncu3.log 2>&1
tail -2 gpurun_out/r02f_ncu2.log gpurun_out/r02f_ncu3.log
