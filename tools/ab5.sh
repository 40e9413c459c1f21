run() { tag=$1; lib=$2; n=$3; shift 3
  env "$@" SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$lib.so SFDTD_VERBOSE=1 python bench.py --steps 2 --warmup 1 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log)"; }
run qt_1 pz_qt 14208 SFDTD_QUEUE=1 SFDTD_QSLICES=1
run qt_8 pz_qt 14208 SFDTD_QUEUE=1
run qt_16 pz_qt 14208 SFDTD_QUEUE=1 SFDTD_QSLICES=16
run qt_8_t1 pz_qt 14208 SFDTD_QUEUE=1 SFDTD_QTAIL=1.0
run qt_8_t2 pz_qt 14208 SFDTD_QUEUE=1 SFDTD_QTAIL=2.0
run qt_8_28k pz_qt 28416 SFDTD_QUEUE=1
SFDTD_QUEUE=1 SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_pz_qt.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_pz_qt.log 2>&1; echo "pytest pz_qt rc=$? $(tail -1 gpurun_out/t_pz_qt.log)"
