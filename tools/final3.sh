bash tools/prof.sh > gpurun_out/prof_sh.log 2>&1; echo "prof rc=$?"
python bench.py > gpurun_out/b_r01e_default.log 2> gpurun_out/b_r01e_default.err; echo "default rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/b_r01e_default.log | head -3 | tr '\n' ' ') $(grep -o '"step_ms": [^]]*]' gpurun_out/b_r01e_default.log)"
