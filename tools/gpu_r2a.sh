#!/bin/bash
# round-2 check A: GPU parity suite, then the headline bench with the pipeline traced, chunk sweep, ascii vs packed
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
for pc in 1 2 3; do
  PLB_PIPE_CHUNKS=$pc python bench.py --steps 10 --warmup 3 --no-cpu 2> gpurun_out/bench_r2a_pc$pc.err | tail -1 > gpurun_out/bench_r2a_pc$pc.json
done
python bench.py --steps 10 --warmup 3 --no-cpu --ascii 2> gpurun_out/bench_r2a_ascii.err | tail -1 > gpurun_out/bench_r2a_ascii.json
PLB_TRACE=1 python bench.py --steps 4 --warmup 3 --no-cpu 2> gpurun_out/bench_r2a_trace.err | tail -1 > gpurun_out/bench_r2a_trace.json
python - <<'PY'
import json
for f in ("pc1", "pc2", "pc3", "ascii", "trace"):
    try:
        d = json.load(open("gpurun_out/bench_r2a_%s.json" % f))
        print(f, "value %.1f GCUPS %.3f ms | e2e %.1f GCUPS %.3f ms h2d %.1f MB single %s each %s" % (
            d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"] / 1e6,
            d["e2e"]["single_call_ms"], d["e2e"]["ms_each_step"]))
        print("   ", {k: round(v, 3) for k, v in d["roofline"]["kernel_ms_all"].items()})
    except Exception as e:
        print(f, "unreadable", e)
PY
grep "plb\]" gpurun_out/bench_r2a_trace.err | tail -24
tail -3 gpurun_out/bench_r2a_pc2.err
