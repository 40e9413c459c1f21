python -m pytest tests -m gpu -x -q > gpurun_out/t_final2.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/t_final2.log)"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final2.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke_final2.log)"
python bench.py > gpurun_out/b_r01d_default.log 2> gpurun_out/b_r01d_default.err; echo "default rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/b_r01d_default.log | head -3 | tr '\n' ' ') $(grep -o '"step_ms": [^]]*]' gpurun_out/b_r01d_default.log)"
python bench.py --strings 14208 --no-cpu-baseline > gpurun_out/b_r01d_14k.log 2> gpurun_out/b_r01d_14k.err; echo "14k rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/b_r01d_14k.log | head -2 | tr '\n' ' ')"
for e in hammer bow random; do python bench.py --steps 2 --warmup 1 --length 0.1 --strings 3552 --excitation $e --no-cpu-baseline --no-e2e > gpurun_out/ex_$e.log 2> gpurun_out/ex_$e.err; echo "$e $(grep -o '"value": [0-9.]*' gpurun_out/ex_$e.log | head -1)"; done
