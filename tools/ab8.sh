run() { tag=$1; lib=$2; n=$3; shift 3
  env "$@" SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$lib.so SFDTD_VERBOSE=1 python bench.py --steps 2 --warmup 1 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log)"; }
for v in base tb16 trim both; do run $v $v 28416; done
SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_both.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_workload.py -m gpu -x -q -k "not mixed and not hammer and not bow" > gpurun_out/t_both.log 2>&1; echo "pytest both rc=$? $(tail -1 gpurun_out/t_both.log)"
