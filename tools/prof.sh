# ncu evidence for profiles/: launch list of one warm-up + one timed step, and a full capture of the largest bucket kernel
set -x
B="python bench.py --steps 1 --warmup 1 --length 0.02 --strings 14208 --no-cpu-baseline --no-e2e"
$B > gpurun_out/prof_plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01e.csv $B > gpurun_out/prof_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 7 --launch-count 1 -f -o gpurun_out/prof_r01e $B > gpurun_out/prof_ncu2.log 2>&1
ls -la gpurun_out/
