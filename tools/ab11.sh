run() { tag=$1; n=$2; shift 2
  env "$@" SFDTD_VERBOSE=1 python bench.py --steps 3 --warmup 2 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log) $(grep -o '"step_ms": [^]]*]' gpurun_out/ab_$tag.log)"; }
for d in 4 8 16; do run ld${d}_28k 28416 SFDTD_LANE_DIV=$d; done
run ld8_14k 14208 SFDTD_LANE_DIV=8
