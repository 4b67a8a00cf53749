#!/bin/bash
# ncu captures of the two big kernels (run under gpurun). usage: tools/gpu_ncu.sh TAG
TAG=${1:-x}
NCU=/usr/local/cuda/bin/ncu
mkdir -p gpurun_out
for K in k_dp k_anchor; do
$NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${K}_$TAG -f \
    python bench.py --steps 1 --warmup 3 --windows 10000 > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
