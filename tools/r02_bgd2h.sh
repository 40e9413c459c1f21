# does a bulk device->host copy beside the stepper slow the stepper down?  (main leg only, 0.5 s strings; GB per step, chunk MiB)
B="python bench.py --steps 3 --warmup 1 --length 0.5 --no-cpu-baseline --no-drop-in --no-dataset --no-grouped --no-fp32 --no-e2e"
P='import sys,json; d=json.loads(sys.stdin.read()); print(d["step_ms"])'
echo "no copy          $($B 2>/dev/null | tail -1 | python -c "$P")"
echo "6 GB, 256 MiB    $(SFDTD_BENCH_BG_D2H=6 $B 2>/dev/null | tail -1 | python -c "$P")"
echo "6 GB, 16 MiB     $(SFDTD_BENCH_BG_D2H=6 SFDTD_BENCH_BG_CHUNK_MB=16 $B 2>/dev/null | tail -1 | python -c "$P")"
echo "no copy          $($B 2>/dev/null | tail -1 | python -c "$P")"
