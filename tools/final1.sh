python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/t_final.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final.log)"
python tools/config_latency.py > gpurun_out/latency_final.log 2>&1; echo "latency rc=$?"; tail -5 gpurun_out/latency_final.log
for e in hammer bow random; do python bench.py --steps 2 --warmup 1 --length 0.1 --strings 3552 --excitation $e --no-cpu-baseline --no-e2e > gpurun_out/ex_$e.log 2> gpurun_out/ex_$e.err; echo "$e $(grep -o '"value": [0-9.]*' gpurun_out/ex_$e.log | head -1)"; done
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/b_r01c_ref.log 2>&1; echo "ref rc=$?"; cut -c1-300 gpurun_out/b_r01c_ref.log | tail -1
