python bench.py > gpurun_out/b_r01c_default.log 2> gpurun_out/b_r01c_default.err; echo "default rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/b_r01c_default.log | head -1)"
python bench.py --strings 28416 --no-cpu-baseline > gpurun_out/b_r01c_28k.log 2> gpurun_out/b_r01c_28k.err; echo "28k rc=$? $(grep -o '"value": [0-9.]*' gpurun_out/b_r01c_28k.log | head -1)"
nvidia-smi --query-gpu=memory.total,memory.used --format=csv
free -g | head -2
