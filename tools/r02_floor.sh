# A/B of sweep-floor variants: speed (pluck, kernel only) + parity errors on the pluck fixtures
#   gpurun -- 'bash tools/r02_floor.sh <tag> base f33 ...'
tag=$1; shift
for v in "$@"; do
  export SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so
  timeout 600 python bench.py --steps 2 --warmup 1 --length 0.2 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped > gpurun_out/${tag}_${v}.log 2> gpurun_out/${tag}_${v}.err
  echo "$v $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_${v}.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/${tag}_${v}.log) $(grep -o '"mean_sweeps_per_step": [0-9.]*' gpurun_out/${tag}_${v}.log)"
  timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -k "pluck or random" > gpurun_out/${tag}_${v}_t.log 2>&1
  echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_${v}_t.log)"
done
