#!/bin/bash
# refresh of the round-2 evidence after the last kernel changes: headline line, launch list, full captures of k_dp / k_anchor
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_r02_n1.err | tail -1 > gpurun_out/bench_r02_n1.json
python bench.py --config 3 --windows 100000 --steps 5 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_r02_config3_100k.json
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_r02_config3.json
python bench.py --stage select --steps 10 --warmup 3 2> /dev/null | tail -1 > gpurun_out/bench_r02_select.json
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_r02.log 2>&1
for K in k_dp k_anchor; do
  $NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${K}_r02_final -f \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${K}_r02.log 2>&1
done
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_r02_config3_100k_launches.csv \
    python bench.py --config 3 --windows 100000 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3_launch.log 2>&1
for K in k_dp k_anchor; do
  $NCU --set full --clock-control none -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${K}_r02_config3_100k -f \
      python bench.py --config 3 --windows 100000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c3_$K.log 2>&1
done
python - <<'PY'
import json
for f in ("n1", "config3", "config3_100k", "select"):
    d = json.load(open("gpurun_out/bench_r02_%s.json" % f))
    e = d.get("e2e", {})
    print(f, "value %.1f %.3f ms/step | e2e %.1f %s" % (d["value"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step")), (d.get("roofline") or {}).get("kernel_ms_all"))
PY
