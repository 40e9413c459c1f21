# ncu evidence for profiles/ (pluck headline workload, 0.02 s per string so that ncu's replays stay short):
# launch list of one warm-up + one timed step, and a full capture of the largest bucket kernel
set -x
B="python bench.py --steps 1 --warmup 1 --length 0.02 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped"
SFDTD_VERBOSE=1 $B > gpurun_out/r02_prof_plain.log 2> gpurun_out/r02_prof_plain.err || exit 1
grep -h "bucket" gpurun_out/r02_prof_plain.err | head -10
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02.csv $B > gpurun_out/r02_prof_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 15 --launch-count 1 -f -o gpurun_out/prof_r02_pluck $B > gpurun_out/r02_prof_ncu2.log 2>&1
tail -2 gpurun_out/r02_prof_ncu2.log
