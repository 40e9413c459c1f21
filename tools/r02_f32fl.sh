tag=$1; shift
for v in "$@"; do
  export SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so
  timeout 600 python bench.py --steps 1 --warmup 1 --fp32-steps 3 --length 0.2 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped > gpurun_out/${tag}_${v}.log 2> gpurun_out/${tag}_${v}.err
  echo "$v fp32 $(grep -o '"fp32": {"dtype": "f32", "value": [0-9.]*' gpurun_out/${tag}_${v}.log) $(grep -o '"mean_sweeps_per_step": [0-9.]*' gpurun_out/${tag}_${v}.log | head -1) $(grep -o '"uout_rel_l2_vs_fp64_first_50ms_p50_p90_p99": [^]]*]' gpurun_out/${tag}_${v}.log)"
  timeout 900 python -m pytest tests/test_gpu_fp32.py -q -s > gpurun_out/${tag}_${v}_t.log 2>&1; echo "$v fp32 tests rc=$? $(tail -1 gpurun_out/${tag}_${v}_t.log)"
  grep -h "ref32\|fp32 vs" gpurun_out/${tag}_${v}_t.log | grep -v "line.append\|print(" | sed 's/; v_r_out.*//' | cut -c1-200
done
