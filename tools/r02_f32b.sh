tag=$1
timeout 900 python -m pytest tests/test_gpu_fp32.py -q -s > gpurun_out/${tag}_f32.log 2>&1; echo "fp32 rc=$? $(tail -1 gpurun_out/${tag}_f32.log)"
grep -h "ref32\|fp32 vs" gpurun_out/${tag}_f32.log
