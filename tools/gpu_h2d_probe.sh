#!/bin/bash
# aggregate H2D bandwidth of the box at N = 1, 2, 4, 8 concurrent ranks (tools/h2d_probe.py)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
lscpu | head -30 > gpurun_out/lscpu.txt 2>&1
(numactl -H || true) >> gpurun_out/lscpu.txt 2>&1
free -g >> gpurun_out/lscpu.txt
: > gpurun_out/h2d_probe.jsonl
for n in 1 2 4 8; do
  timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
     tools/h2d_probe.py 2>/dev/null | grep '^{' >> gpurun_out/h2d_probe.jsonl
done
cat gpurun_out/h2d_probe.jsonl
