tag=$1
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_t.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_t.log)"
grep -h "^FAILED\|^ERROR" gpurun_out/${tag}_t.log | head
grep -h "pluck_b24_1s\|uout\[\|zout\[" gpurun_out/${tag}_t.log | grep "nan\|worst" | sort -k7 | tail -8
