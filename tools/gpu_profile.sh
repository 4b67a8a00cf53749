#!/bin/bash
# Runs on the GPU box (under gpurun): microbenchmark, both bench arms, ncu launch list and one
# full ncu capture of the dominant kernel.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
./tools/microbench > gpurun_out/microbench_$TAG.txt 2>&1
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -c 3000 gpurun_out/bench_$TAG.err
NCU=/usr/local/cuda/bin/ncu
$NCU --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --windows 10000 > gpurun_out/ncu_launch_$TAG.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:k_dp -s 3 -c 1 -o gpurun_out/prof_kdp_$TAG -f \
    python bench.py --steps 1 --warmup 3 --windows 10000 > gpurun_out/ncu_full_$TAG.log 2>&1
$NCU --set full --clock-control none --import-source on -k regex:k_anchor -s 3 -c 1 -o gpurun_out/prof_kanchor_$TAG -f \
    python bench.py --steps 1 --warmup 3 --windows 10000 > gpurun_out/ncu_full_anchor_$TAG.log 2>&1
ls -la gpurun_out
cat gpurun_out/microbench_$TAG.txt
cat gpurun_out/bench_ref_$TAG.json
cat gpurun_out/bench_$TAG.json
