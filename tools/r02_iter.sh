# one development iteration on the GPU box: test-suite, then short bench lines of every excitation
#   gpurun -- 'bash tools/r02_iter.sh <tag> [pluck_strings]'
tag=$1; np=${2:-28416}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_t.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_t.log)"
for ex in hammer bow random; do
  SFDTD_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 1 --length 0.1 --strings 3552 --excitation $ex --no-cpu-baseline --no-e2e > gpurun_out/${tag}_$ex.log 2> gpurun_out/${tag}_$ex.err
  echo "$ex rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_$ex.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/${tag}_$ex.log) $(grep -o '"health": {[^}]*}' gpurun_out/${tag}_$ex.log)"
done
timeout 900 python bench.py --steps 3 --warmup 2 --length 0.2 --strings $np --no-cpu-baseline --no-e2e > gpurun_out/${tag}_pluck.log 2> gpurun_out/${tag}_pluck.err
echo "pluck rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_pluck.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/${tag}_pluck.log)"
grep -h "bucket\|timing" gpurun_out/${tag}_hammer.err | tail -8
tail -5 gpurun_out/${tag}_t.log
