run() { tag=$1; shift
  env "$@" python bench.py --steps 3 --warmup 2 --length 0.2 --no-cpu-baseline --no-e2e > gpurun_out/sw_$tag.log 2> gpurun_out/sw_$tag.err
  echo "$tag $(grep -o '"step_ms": [^]]*]' gpurun_out/sw_$tag.log)"; }
run def
run t05 SFDTD_QTAIL=0.5
run t2 SFDTD_QTAIL=2.0
run s16 SFDTD_QSLICES=16
run s4 SFDTD_QSLICES=4
