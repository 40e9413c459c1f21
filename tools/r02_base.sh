# round-2 baseline: GPU test-suite, short bench lines of every excitation, ncu launch list + full capture of the grouped kernel
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_t_base.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/r02_t_base.log)"
python bench.py --steps 3 --warmup 2 --length 0.2 --strings 28416 --no-cpu-baseline --no-e2e > gpurun_out/r02_base_pluck.log 2> gpurun_out/r02_base_pluck.err
for ex in hammer bow random; do
  python bench.py --steps 2 --warmup 1 --length 0.1 --strings 3552 --excitation $ex --no-cpu-baseline --no-e2e > gpurun_out/r02_base_$ex.log 2> gpurun_out/r02_base_$ex.err
done
grep -o '"value": [0-9.]*' gpurun_out/r02_base_*.log
B="python bench.py --steps 1 --warmup 1 --length 0.01 --strings 3552 --excitation hammer --no-cpu-baseline --no-e2e"
SFDTD_VERBOSE=1 $B > gpurun_out/r02_prof_plain.log 2> gpurun_out/r02_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches_r02a_hammer.csv $B > gpurun_out/r02_ncu1.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:step_kernel --launch-skip 1 --launch-count 1 -f -o gpurun_out/prof_r02a_hammer $B > gpurun_out/r02_ncu2.log 2>&1
ls -la gpurun_out/*r02*
