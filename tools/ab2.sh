# A/B harness: bench every variant library in torch_fdtd_string_b200/ab named on the command line (tier 2), then the GPU
# test-suite against the last two
B="python bench.py --steps 2 --warmup 1 --length 0.2 --strings 14208 --no-cpu-baseline --no-e2e"
for v in "$@"; do
  SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so SFDTD_VERBOSE=1 $B > gpurun_out/ab_$v.log 2> gpurun_out/ab_$v.err; echo "$v rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$v.log)"
done
for v in "${@: -2}"; do
  SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so python -m pytest tests -m gpu -x -q > gpurun_out/t_$v.log 2>&1; echo "pytest $v rc=$? $(tail -1 gpurun_out/t_$v.log)"
done
