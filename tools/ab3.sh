# usage: ab3.sh <tag> <lib> <strings> [ENV=VAL ...]   -- one bench run of a variant library
run() { tag=$1; lib=$2; n=$3; shift 3
  env "$@" SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$lib.so SFDTD_VERBOSE=1 python bench.py --steps 2 --warmup 1 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log)"; }
run pz_l pz_l 14208
run pz_q0 pz_q 14208 SFDTD_QUEUE=0
run pz_q1 pz_q 14208 SFDTD_QUEUE=1
run pz_q0_28k pz_q 28416 SFDTD_QUEUE=0
run pz_q1_28k pz_q 28416 SFDTD_QUEUE=1
SFDTD_QUEUE=1 SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_pz_q.so python -m pytest tests -m gpu -x -q > gpurun_out/t_pz_q1.log 2>&1; echo "pytest pz_q1 rc=$? $(tail -1 gpurun_out/t_pz_q1.log)"
