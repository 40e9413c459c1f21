# A/B harness (GPU box): bench every variant library torch_fdtd_string_b200/ab/lib_<name>.so named on the command line twice,
# interleaved (28416 strings x 0.2 s, kernel only), then run the GPU test-suite against the last one.
#   nvcc ... -D<variant flags> -o torch_fdtd_string_b200/ab/lib_new.so torch_fdtd_string_b200/csrc/sfdtd.cu
#   gpurun -- 'bash tools/ab.sh base new'
# Compare the per-step times (step_ms), not single values: the first timed step after a short warm-up can be an outlier.
run() { tag=$1; lib=$2; n=$3; shift 3
  env "$@" SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$lib.so python bench.py --steps 3 --warmup 2 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log) $(grep -o '"step_ms": [^]]*]' gpurun_out/ab_$tag.log)"; }
V="$@"
for r in 1 2; do for v in $V; do run ${v}_$r $v 28416; done; done
last=${@: -1}
SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$last.so timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_$last.log 2>&1; echo "pytest $last rc=$? $(tail -1 gpurun_out/t_$last.log)"
