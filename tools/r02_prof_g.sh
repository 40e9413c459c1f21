# ncu --set full of the grouped (cluster) kernel on the hammer workload; $1 = tag, $2 = launches to skip
tag=$1; skip=${2:-5}
B="python bench.py --steps 1 --warmup 1 --length 0.01 --strings 3552 --excitation hammer --no-cpu-baseline --no-e2e"
SFDTD_VERBOSE=1 $B > gpurun_out/${tag}_plain.log 2> gpurun_out/${tag}_plain.err || exit 1
grep -h "bucket\|timing" gpurun_out/${tag}_plain.err | tail -8
ncu --set full --import-source on --clock-control none -k regex:group_kernel --launch-skip $skip --launch-count 1 -f -o gpurun_out/prof_${tag} $B > gpurun_out/${tag}_ncu.log 2>&1
tail -3 gpurun_out/${tag}_ncu.log
