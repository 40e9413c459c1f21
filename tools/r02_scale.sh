# multi-GPU records: $1 = N (GPUs of this box).  Default bench (kernel + e2e) and the configs[4] sweep, torchrun like the driver.
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
$TR bench.py --gpus $N --steps 3 --warmup 3 --no-drop-in --no-dataset --no-grouped > gpurun_out/r02_scale_n$N.log 2> gpurun_out/r02_scale_n$N.err
tail -c 1800 gpurun_out/r02_scale_n$N.log
$TR bench.py --gpus $N --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-drop-in --no-dataset --no-grouped --sweep ${2:-1008,4080,16368,65520,262128,1048560} > gpurun_out/r02_sweep_n$N.log 2> gpurun_out/r02_sweep_n$N.err
grep "\[sweep\]" gpurun_out/r02_sweep_n$N.err
