#!/bin/bash
# round-2 run F (8 GPUs): headline bench at N=8 with the pipeline traced, BASELINE config 4 at size, config 5 at 1/2/4/8
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 "${@:2}"; }
PLB_TRACE=1 run 8 --steps 10 --warmup 3 2> gpurun_out/bench_r2f_n8.err | tail -1 > gpurun_out/bench_r2f_n8.json
run 8 --config 4 --steps 5 --warmup 3 2> gpurun_out/bench_r2f_c4.err | tail -1 > gpurun_out/bench_r2f_c4.json
for n in 8 4 2; do
  run $n --config 5 --no-cpu 2> gpurun_out/bench_r2f_c5_n$n.err | tail -1 > gpurun_out/bench_r2f_c5_n$n.json
done
python bench.py --config 5 2> gpurun_out/bench_r2f_c5_n1.err | tail -1 > gpurun_out/bench_r2f_c5_n1.json
python - <<'PY'
import json
for f in ("n8", "c4", "c5_n1", "c5_n2", "c5_n4", "c5_n8"):
    try:
        d = json.load(open("gpurun_out/bench_r2f_%s.json" % f))
        print(f, "value %.1f GCUPS %.3f ms/step | e2e %.1f %s | total_ms %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("total_ms")))
        print("   ", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_all"].items()}, d.get("oracle_check"), d.get("gather_check"))
        if f == "n8": print("    e2e each", d["e2e"]["ms_each_step"], "single", d["e2e"]["single_call_ms"])
    except Exception as e:
        print(f, "unreadable", e)
PY
for f in n8 c4 c5_n8 c5_n1; do echo "== $f"; grep -v "^\[plb\]\|Setting OMP\|^\*\*\*\|^$\|k:[0-9]" gpurun_out/bench_r2f_$f.err | tail -4; done
grep "plb\] job" gpurun_out/bench_r2f_n8.err | tail -12
