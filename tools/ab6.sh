run() { tag=$1; lib=$2; exc=$3; n=$4; len=$5; shift 5
  env "$@" SFDTD_LIB=$lib SFDTD_VERBOSE=1 python bench.py --steps 2 --warmup 1 --length $len --strings $n --excitation $exc --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log)"; }
NEW=$PWD/torch_fdtd_string_b200/libsfdtd.so; BASE=$PWD/torch_fdtd_string_b200/ab/lib_base.so
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_new.log 2>&1; echo "pytest new rc=$? $(tail -1 gpurun_out/t_new.log)"
run new_pluck $NEW pluck 14208 0.2
for e in hammer bow random; do run base_$e $BASE $e 3552 0.1; run new_$e $NEW $e 3552 0.1; done
