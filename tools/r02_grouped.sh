# grouped-mode bench lines: $1 = tag, $2 = strings, $3 = length
tag=$1; n=${2:-14208}; len=${3:-0.05}
for ex in hammer bow random; do
  SFDTD_VERBOSE=1 timeout 600 python bench.py --steps 2 --warmup 1 --length $len --strings $n --excitation $ex --no-cpu-baseline --no-e2e > gpurun_out/${tag}_${ex}_$n.log 2> gpurun_out/${tag}_${ex}_$n.err
  echo "$ex $n rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/${tag}_${ex}_$n.log | head -1) $(grep -o '"frac": [0-9.]*' gpurun_out/${tag}_${ex}_$n.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/${tag}_${ex}_$n.log)"
done
grep -h "bucket\|timing" gpurun_out/${tag}_hammer_$n.err | tail -10
