#!/bin/bash
# round-2 check E (2 GPUs): the multi-GPU paths at small sizes before the 8-GPU run
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $1 "${@:2}"; }
PLB_TRACE=1 run 2 --steps 6 --warmup 3 2> gpurun_out/bench_r2e_n2.err | tail -1 > gpurun_out/bench_r2e_n2.json
run 2 --config 4 --windows 3000 --steps 5 --warmup 3 2> gpurun_out/bench_r2e_c4.err | tail -1 > gpurun_out/bench_r2e_c4.json
run 2 --config 5 --c5-windows 1184 2> gpurun_out/bench_r2e_c5.err | tail -1 > gpurun_out/bench_r2e_c5.json
python bench.py --config 5 --c5-windows 592 2> gpurun_out/bench_r2e_c5n1.err | tail -1 > gpurun_out/bench_r2e_c5n1.json
for f in n2 c4 c5 c5n1; do echo "== $f"; grep -v "^\[plb\]\|Setting OMP\|^\*\*\*\|^$" gpurun_out/bench_r2e_$f.err | tail -4; done
python - <<'PY'
import json
for f in ("n2", "c4", "c5", "c5n1"):
    try:
        d = json.load(open("gpurun_out/bench_r2e_%s.json" % f))
        print(f, "value %.1f GCUPS %.3f ms/step | e2e %.1f %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step")))
        print("   ", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_all"].items()}, d.get("oracle_check"), d.get("gather_check"), d.get("total_ms"))
    except Exception as e:
        print(f, "unreadable", e)
PY
grep "plb\] job" gpurun_out/bench_r2e_n2.err | tail -6
