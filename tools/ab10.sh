run() { tag=$1; n=$2; shift 2
  env "$@" SFDTD_VERBOSE=1 python bench.py --steps 3 --warmup 2 --length 0.2 --strings $n --no-cpu-baseline --no-e2e > gpurun_out/ab_$tag.log 2> gpurun_out/ab_$tag.err
  echo "$tag $(grep -o '"value": [0-9.]*' gpurun_out/ab_$tag.log) $(grep -o '"step_ms": [^]]*]' gpurun_out/ab_$tag.log)"; }
for w in 16 32 64; do run wl${w}_28k 28416 SFDTD_WLMIN=$w; done
for w in 16 64; do run wl${w}_14k 14208 SFDTD_WLMIN=$w; done
