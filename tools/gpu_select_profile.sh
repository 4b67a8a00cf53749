#!/bin/bash
# N1 on the GPU box: smoke(), launch list of one selection call, full ncu captures of its own kernels and of the
# scoring kernels in the selection shape (a late round).  usage: tools/gpu_select_profile.sh TAG
TAG=${1:-r01s}
NCU=/usr/local/cuda/bin/ncu
[ -z "$SKIP_SMOKE" ] && python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 $NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_select_$TAG.csv \
    python tools/select_bench.py 10000 1 --pin > gpurun_out/ncu_launch_select_$TAG.log 2>&1
for K in k_build_haps k_trial_score k_anchor k_dp; do
timeout 300 $NCU --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -o gpurun_out/prof_select_${K}_$TAG -f \
    python tools/select_bench.py 10000 1 --pin > gpurun_out/ncu_select_${K}_$TAG.log 2>&1
done
ls -la gpurun_out/*select*$TAG* | head -20
