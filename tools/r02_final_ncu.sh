# ncu evidence of round 2 (run after `python bench.py` exited 0 in the same or an earlier call): launch list of this library's
# kernels, full captures of the largest bucket kernel of the fp64 and of the fp32 build
set -x
B="python bench.py --steps 1 --warmup 1 --fp32-steps 1 --length 0.02 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped"
$B > gpurun_out/r02f_prof_plain.log 2> gpurun_out/r02f_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:sfdtd -c 200 --csv --log-file gpurun_out/launches_r02f.csv $B > gpurun_out/r02f_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"step_kernel<double, .int.16, .int.4" --launch-skip 5 --launch-count 1 -f -o gpurun_out/prof_r02f_f64 $B > gpurun_out/r02f_ncu2.log 2>&1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"step_kernel<float, .int.16, .int.4" --launch-skip 5 --launch-count 1 -f -o gpurun_out/prof_r02f_f32 $B > gpurun_out/r02f_ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
