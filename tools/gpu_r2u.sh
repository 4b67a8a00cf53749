#!/bin/bash
# round-2 run U (8 GPUs): headline bench at N=8 with 2-bit bases + 6-bit quality codes, then the same with the pipeline traced
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 "${@:2}"; }
run 8 --steps 20 --warmup 3 2> gpurun_out/bench_r2u_n8.err | tail -1 > gpurun_out/bench_r2u_n8.json
PLB_TRACE=1 run 8 --steps 8 --warmup 3 2> gpurun_out/trace_r2u_n8.txt | tail -1 > gpurun_out/bench_r2u_n8_traced.json
run 4 --steps 20 --warmup 3 2> gpurun_out/bench_r2u_n4.err | tail -1 > gpurun_out/bench_r2u_n4.json
python - <<'PY'
import json
for f in ("n8", "n8_traced", "n4"):
    try:
        d = json.load(open("gpurun_out/bench_r2u_%s.json" % f))
        print(f, "value %.1f GCUPS %.3f ms/step | e2e %.1f %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step")))
        print("    e2e each", d["e2e"]["ms_each_step"], "single", d["e2e"]["single_call_ms"], d["e2e"]["h2d_bytes_per_step"], d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
grep -v "^\[plb\]\|Setting OMP\|^\*\*\*\|^$\|k:[0-9]" gpurun_out/bench_r2u_n8.err | tail -4
grep "plb\] job" gpurun_out/trace_r2u_n8.txt | tail -10
