#!/bin/bash
# last pass of a round: whole GPU parity suite, select-stage bench (with its CPU leg), headline bench
TAG=${1:-r01v}
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
PLB_TRACE=1 python bench.py --stage select --steps 10 --warmup 3 2>gpurun_out/bench_select_$TAG.err | tail -1 > gpurun_out/bench_select_$TAG.json
tail -2 gpurun_out/bench_select_$TAG.err
python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/bench_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/bench_select_$TAG.json")); print("select value %.1f ms %.2f e2e %.1f ms %.2f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]), d.get("cpu_baseline",{}).get("windows_per_s"))
d=json.load(open("gpurun_out/bench_$TAG.json")); print("path value %.1f ms %.3f e2e %.1f ms %.3f"%(d["value"],d["ms_per_step"],d["e2e"]["value"],d["e2e"]["ms_per_step"]))
PY
