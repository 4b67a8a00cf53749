#!/bin/bash
# quick GPU check: parity tests + one bench line summarised
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python bench.py --steps ${1:-5} --warmup 3 2> gpurun_out/bench_quick.err | tail -1 > gpurun_out/bench_quick.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_quick.json"))
print("value %.1f GCUPS  %.3f ms/step   e2e %.1f GCUPS %.3f ms" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v, 3) for k, v in d["roofline"]["kernel_ms_all"].items()})
print(d["stats"], d["clocks"])
PY
tail -5 gpurun_out/bench_quick.err
