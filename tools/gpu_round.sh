#!/bin/bash
# end-of-milestone GPU pass: whole parity suite, headline bench (+ reference arm), select-stage bench, launch lists
tag=${1:-r01s}
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_${tag}.err | tail -1 > gpurun_out/bench_${tag}.json
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_ref_${tag}.err | tail -1 > gpurun_out/bench_ref_${tag}.json
python bench.py --stage select --steps 10 --warmup 3 2> gpurun_out/bench_select_${tag}.err | tail -1 > gpurun_out/bench_select_${tag}.json
python bench.py --stage select --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_select_ref_${tag}.err | tail -1 > gpurun_out/bench_select_ref_${tag}.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_select_${tag}.csv \
    python bench.py --stage select --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_launch_select_${tag}.log 2>&1
python - <<PY
import json
for f in ("bench_${tag}", "bench_ref_${tag}", "bench_select_${tag}", "bench_select_ref_${tag}"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f))
        print(f, "value %.1f ms %.3f e2e %.1f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]), d.get("clocks"))
    except Exception as e:
        print(f, "unreadable", e)
PY
