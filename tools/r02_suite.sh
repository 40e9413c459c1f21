tag=$1
timeout 2400 python -m pytest tests -m gpu -q -s > gpurun_out/${tag}_t.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/${tag}_t.log)"
grep -h "^FAILED\|^ERROR" gpurun_out/${tag}_t.log | head
grep -h "lowf0" gpurun_out/${tag}_t.log | cut -c1-300 | head
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$? $(tail -3 gpurun_out/${tag}_smoke.log | tr '\n' ' ')"
