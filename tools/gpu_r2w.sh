#!/bin/bash
# round-2 run W (8 GPUs): BASELINE config 4 with the final tree (2-bit bases + 6-bit quality codes in the e2e leg)
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus $1 "${@:2}"; }
run 8 --config 4 --steps 5 --warmup 3 2> gpurun_out/bench_r2w_c4.err | tail -1 > gpurun_out/bench_r2w_c4.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2w_c4.json"))
print("c4 value %.1f GCUPS %.3f ms/step | e2e %.1f %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step")), d.get("gather_check"), d["e2e"]["h2d_bytes_per_step"])
PY
tail -3 gpurun_out/bench_r2w_c4.err
