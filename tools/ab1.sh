set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r01b.log 2>&1; echo "pytest rc=$?"
B="python bench.py --steps 2 --warmup 1 --length 0.2 --strings 14208 --no-cpu-baseline --no-e2e"
for v in base pred; do for t in 2 3; do
  SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so SFDTD_TIER=$t SFDTD_VERBOSE=1 $B > gpurun_out/ab_${v}_t$t.log 2> gpurun_out/ab_${v}_t$t.err; echo "$v t$t rc=$?"
done; done
grep -o '"value": [0-9.]*' gpurun_out/ab_*.log
