run() { tag=$1; n=$2; len=$3; exc=$4; shift 4
  env "$@" SFDTD_VERBOSE=1 python bench.py --steps 2 --warmup 1 --length $len --strings $n --excitation $exc --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log)"; }
run t2 14208 0.2 pluck SFDTD_TIER=2
run t3 14208 0.2 pluck SFDTD_TIER=3
run t0 14208 0.2 pluck SFDTD_TIER=0
run ham 3552 0.1 hammer
SFDTD_TIER=3 timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/t_t3.log 2>&1; echo "pytest t3 rc=$? $(tail -1 gpurun_out/t_t3.log)"
