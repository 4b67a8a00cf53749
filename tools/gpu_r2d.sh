#!/bin/bash
# round-2 check D (1 GPU): parity suite, config 5 on a small job, headline bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py --config 5 --c5-windows 128 --c5-chunk 32 2> gpurun_out/bench_r2d_c5.err | tail -1 > gpurun_out/bench_r2d_c5.json
tail -5 gpurun_out/bench_r2d_c5.err
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_r2d.err | tail -1 > gpurun_out/bench_r2d.json
tail -3 gpurun_out/bench_r2d.err
python - <<'PY'
import json
for f in ("bench_r2d_c5", "bench_r2d"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.1f GCUPS %.3f ms/step | e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
        print("   ", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_all"].items()}, d.get("oracle_check"), d.get("cpu_baseline", {}).get("value") if d.get("cpu_baseline") else None)
    except Exception as e:
        print(f, "unreadable", e)
PY
