set -x
B="python bench.py --steps 1 --warmup 1 --fp32-steps 1 --length 0.02 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped"
$B > gpurun_out/r02g_prof_plain.log 2> gpurun_out/r02g_prof_plain.err || exit 1
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:"step_kernel<float, .int.16, .int.4" --launch-skip 5 --launch-count 1 -f -o gpurun_out/prof_r02g_f32 $B > gpurun_out/r02g_ncu3.log 2>&1
ls -la gpurun_out/prof_r02g_f32.ncu-rep
