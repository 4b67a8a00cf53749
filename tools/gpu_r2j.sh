#!/bin/bash
# quick: headline bench (packed with / without quality codes), stats, k_anchor source profile
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-cpu 2> gpurun_out/bench_r2j.err | tail -1 > gpurun_out/bench_r2j.json
/usr/local/cuda/bin/ncu --set full --clock-control none --import-source on -k regex:k_anchor -s 3 -c 1 -o gpurun_out/prof_k_anchor_r02b -f \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_k_anchor_r02b.log 2>&1
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2j.json"))
print("value %.1f GCUPS %.3f ms | e2e %.1f GCUPS %.3f ms each %s single %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["ms_each_step"], d["e2e"]["single_call_ms"]))
print({k: round(v, 3) for k, v in d["roofline"]["kernel_ms_all"].items()}, d["stats"])
PY
