# fp32 kernels: tests with printed errors, full GPU suite (regression of the templated fp64 build), short bench with the fp32 leg
tag=$1
timeout 900 python -m pytest tests/test_gpu_fp32.py -x -q -s > gpurun_out/${tag}_f32.log 2>&1; echo "fp32 rc=$? $(tail -1 gpurun_out/${tag}_f32.log)"
grep -h "ref32\|fp32 vs" gpurun_out/${tag}_f32.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_t.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_t.log)"
timeout 900 python bench.py --steps 2 --warmup 1 --length 0.2 --strings 28416 --no-cpu-baseline --no-e2e --no-drop-in --no-dataset --no-grouped > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
echo "bench rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_bench.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/${tag}_bench.log)"
grep -o '"fp32": {.*"drop_in"' gpurun_out/${tag}_bench.log | cut -c1-1500
tail -5 gpurun_out/${tag}_bench.err
