run() { tag=$1; lib=$2; n=$3; shift 3
  env "$@" SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$lib.so SFDTD_VERBOSE=1 python bench.py --steps 2 --warmup 1 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log)"; }
SFDTD_QUEUE=1 SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_pz_qs.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_pz_qs.log 2>&1; echo "pytest pz_qs rc=$? $(tail -1 gpurun_out/t_pz_qs.log)"
run qs_def pz_qs 14208 SFDTD_QUEUE=1
run qs_1 pz_qs 14208 SFDTD_QUEUE=1 SFDTD_QSLICES=1
run qs_4 pz_qs 14208 SFDTD_QUEUE=1 SFDTD_QSLICES=4
run qs_18 pz_qs 14208 SFDTD_QUEUE=1 SFDTD_QSLICES=18
run qs_def_28k pz_qs 28416 SFDTD_QUEUE=1
