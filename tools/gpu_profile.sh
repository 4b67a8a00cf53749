#!/bin/bash
# Runs on the GPU box (under gpurun): parity tests, microbenchmark, both bench arms, ncu launch list and
# full ncu captures of the two big kernels.  Everything lands in gpurun_out/.
# usage: tools/gpu_profile.sh TAG
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu_$TAG.txt
[ -x tools/microbench ] && ./tools/microbench > gpurun_out/microbench_$TAG.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python bench.py --steps 5 --warmup 3 --config 3 > gpurun_out/bench_c3_$TAG.json 2> gpurun_out/bench_c3_$TAG.err
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --windows 10000 > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in k_dp k_anchor; do
$NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${K}_$TAG -f \
    python bench.py --steps 1 --warmup 3 --windows 10000 > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
cat gpurun_out/pytest_gpu_$TAG.txt
tail -c 600 gpurun_out/bench_$TAG.err
cat gpurun_out/bench_ref_$TAG.json
cat gpurun_out/bench_$TAG.json
cat gpurun_out/bench_c3_$TAG.json
