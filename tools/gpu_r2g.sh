#!/bin/bash
# round-2 run G (1 GPU): parity suite with the large differential tests; ncu captures (config 2 with source, config 3 at 100k windows)
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
( time timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -v "^$" | tail -12 ) 2>&1 | tail -16
for K in k_anchor k_dp; do
  $NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${K}_r02a -f \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${K}_r02a.log 2>&1
done
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_r02_config3_100k_launches.csv \
    python bench.py --config 3 --windows 100000 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3_launch.log 2>&1
$NCU --set full --clock-control none -k regex:k_dp -s 1 -c 1 -o gpurun_out/prof_k_dp_r02_config3_100k -f \
    python bench.py --config 3 --windows 100000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c3_kdp.log 2>&1
$NCU --set full --clock-control none -k regex:k_anchor -s 1 -c 1 -o gpurun_out/prof_k_anchor_r02_config3_100k -f \
    python bench.py --config 3 --windows 100000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c3_kanchor.log 2>&1
python bench.py --config 3 --windows 100000 --steps 5 --warmup 3 --no-cpu 2> gpurun_out/bench_r02_config3_100k.err | tail -1 > gpurun_out/bench_r02_config3_100k.json
ls -la gpurun_out/*.ncu-rep; tail -2 gpurun_out/ncu_c3_kdp.log; head -c 600 gpurun_out/bench_r02_config3_100k.json
