tag=$1
for v in main t3; do
  if [ $v = main ]; then unset SFDTD_LIB; else export SFDTD_LIB=$PWD/torch_fdtd_string_b200/ab/lib_$v.so; fi
  timeout 900 python -m pytest tests/test_gpu_fp32.py -q -s > gpurun_out/${tag}_${v}_f32.log 2>&1; echo "$v fp32 rc=$? $(tail -1 gpurun_out/${tag}_${v}_f32.log)"
  grep -h "ref32\|fp32 vs" gpurun_out/${tag}_${v}_f32.log | grep -v line.append | cut -c1-250
done
unset SFDTD_LIB
timeout 900 python -m pytest tests/test_gpu_parity.py -q -s -k "nsynth_batch or single_string_full or golden" > gpurun_out/${tag}_par.log 2>&1; echo "parity rc=$? $(tail -1 gpurun_out/${tag}_par.log)"
grep -h "worst" gpurun_out/${tag}_par.log | sort -t' ' -k12 | awk '{ if ($NF+0 > 0.05) print }' | head -20
