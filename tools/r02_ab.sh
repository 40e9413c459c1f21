# A/B of library variants torch_fdtd_string_b200/ab/lib_<name>.so on the pluck workload (kernel only), interleaved twice
#   gpurun -- 'bash tools/r02_ab.sh <tag> base new ...'
tag=$1; shift
for r in 1 2; do for v in "$@"; do
  SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so timeout 600 python bench.py --steps 3 --warmup 2 --length 0.2 --strings 28416 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_ab_${v}_$r.log 2> gpurun_out/${tag}_ab_${v}_$r.err
  echo "$v $r $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_ab_${v}_$r.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/${tag}_ab_${v}_$r.log)"
done; done
