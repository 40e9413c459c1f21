run() { tag=$1; lib=$2; n=$3; shift 3
  env "$@" SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$lib.so python bench.py --steps 3 --warmup 1 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log) $(grep -o '"step_ms": [^]]*]' gpurun_out/ab_$tag.log)"; }
for r in 1 2; do for v in base tb16 both; do run ${v}_28k_$r $v 28416; done; done
for v in base tb16 both; do run ${v}_14k $v 14208; done
