# grouped mode: lanes per string chosen from the longitudinal grid (SFDTD_LANE_DIV), hammer / bow, kernel only
for ex in hammer bow; do for ld in 4 8 16; do
  SFDTD_VERBOSE=1 SFDTD_LANE_DIV=$ld timeout 300 python bench.py --excitation $ex --strings 14208 --length 0.05 --steps 2 --warmup 1 --no-e2e --no-drop-in --no-dataset --no-grouped --no-fp32 --no-cpu-baseline > gpurun_out/ld_${ex}_$ld.log 2> gpurun_out/ld_${ex}_$ld.err
  echo "$ex lane_div=$ld $(grep -o '"value": [0-9.]*' gpurun_out/ld_${ex}_$ld.log | head -1) $(grep -o '"step_ms": [^]]*]' gpurun_out/ld_${ex}_$ld.log) $(grep -h 'bucket grouped' gpurun_out/ld_${ex}_$ld.err | sort -u | head -4 | tr '\n' ';' | cut -c1-400)"
done; done
