# where does the end-to-end leg lose time: stepper time inside the e2e pipeline with / without the read-back and the overlapped preparation
B="python bench.py --steps 2 --warmup 1 --length 0.5 --no-cpu-baseline --no-drop-in --no-dataset --no-grouped --no-fp32"
P='import sys,json; d=json.loads(sys.stdin.read()); e=d["e2e"]; print(round(d["ms_per_step"],1), round(e["ms_per_step"],1), round(e["ms_stepper"],1), round(e["ms_postprocess"],1))'
echo "default      $($B 2>/dev/null | tail -1 | python -c "$P")"
echo "no D2H       $(SFDTD_BENCH_NO_D2H=1 $B 2>/dev/null | tail -1 | python -c "$P")"
# (r02: with the next plan prepared on a side stream beside the running call the stepper took 1554 instead of 1480 ms and a step
#  1626 ms; prepared in the gap after the call: 1493 / 1537 ms; before the prepass results went through mapped pinned memory its
#  small read-back queued behind the bulk PCM read-back: 1641 ms)
