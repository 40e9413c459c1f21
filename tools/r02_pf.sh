# A/B: L2 prefetch of the state-history rows (SAVE_STATE): drop-in leg (latency-bound) and the headline kernel (must not regress)
for v in "$@"; do
  export SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so
  python bench.py --steps 1 --warmup 1 --strings 3552 --no-e2e --no-fp32 --no-grouped --no-dataset --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v drop-in', d['drop_in']['sequential'], d['drop_in']['in_flight']['value'])"
  python bench.py --steps 2 --warmup 1 --length 0.2 --no-e2e --no-fp32 --no-grouped --no-dataset --no-cpu-baseline --no-drop-in 2>/dev/null | tail -1 | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('$v pluck', d['value'], d['step_ms'])"
done
