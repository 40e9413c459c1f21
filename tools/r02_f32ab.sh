# A/B of fp32 build variants: the fp32 leg of the bench (EXC=pluck|hammer|bow|random, default pluck)
tag=$1; shift
exc=${EXC:-pluck}
if [ $exc = pluck ]; then n=28416; len=0.2; else n=3552; len=0.1; fi
for v in "$@"; do
  export SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so
  timeout 600 python bench.py --steps ${STEPS:-1} --warmup 1 --fp32-steps 3 --length $len --strings $n --excitation $exc --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped > gpurun_out/${tag}_${v}_$exc.log 2> gpurun_out/${tag}_${v}_$exc.err
  echo "$v $exc fp64 $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_${v}_$exc.log | head -1) fp32 $(grep -o '"fp32": {"dtype": "f32", "value": [0-9.]*' gpurun_out/${tag}_${v}_$exc.log) $(grep -o '"mean_sweeps_per_step": [0-9.]*' gpurun_out/${tag}_${v}_$exc.log | head -1)"
done
