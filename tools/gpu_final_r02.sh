#!/bin/bash
# round-2 closing run (1 GPU): smoke(), whole GPU parity suite, both bench arms, config 3, modes, selection stage, then the
# profiler passes: launch lists and full captures of k_dp / k_anchor for config 2 (with source) and config 3 at 100k windows.
# Everything lands in gpurun_out/; the summaries are made from it by tools/ncu_summary.py and copied to profiles/.
TAG=${1:-r02}
mkdir -p gpurun_out
NCU=/usr/local/cuda/bin/ncu
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
( time timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 ) 2>&1 | tail -6 | tee gpurun_out/pytest_gpu_$TAG.txt
python bench.py --impl reference --steps 2 --warmup 1 2> gpurun_out/bench_${TAG}_reference.err | tail -1 > gpurun_out/bench_${TAG}_reference.json
python bench.py --steps 20 --warmup 3 2> gpurun_out/bench_${TAG}_n1.err | tail -1 > gpurun_out/bench_${TAG}_n1.json
python bench.py --steps 10 --warmup 3 --ascii --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_${TAG}_n1_ascii.json
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_${TAG}_config3.json
python bench.py --config 3 --windows 100000 --steps 5 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_${TAG}_config3_100k.json
python bench.py --mode hla --steps 5 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_${TAG}_mode_hla.json
python bench.py --mode flank --steps 3 --warmup 3 --no-cpu 2> /dev/null | tail -1 > gpurun_out/bench_${TAG}_mode_flank.json
python bench.py --stage select --steps 10 --warmup 3 2> /dev/null | tail -1 > gpurun_out/bench_${TAG}_select.json
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_launch_$TAG.log 2>&1
for K in k_dp k_anchor; do
  $NCU --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -o gpurun_out/prof_${K}_${TAG}_final -f \
      python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_${K}_$TAG.log 2>&1
done
$NCU --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/ncu_${TAG}_config3_100k_launches.csv \
    python bench.py --config 3 --windows 100000 --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_c3_launch.log 2>&1
for K in k_dp k_anchor; do
  $NCU --set full --clock-control none -k regex:$K -s 1 -c 1 -o gpurun_out/prof_${K}_${TAG}_config3_100k -f \
      python bench.py --config 3 --windows 100000 --steps 1 --warmup 1 --no-cpu > gpurun_out/ncu_c3_$K.log 2>&1
done
python - <<PY
import json
for f in ("reference", "n1", "n1_ascii", "config3", "config3_100k", "mode_hla", "mode_flank", "select"):
    try:
        d = json.load(open("gpurun_out/bench_${TAG}_%s.json" % f))
        e = d.get("e2e", {})
        print(f, "value %.1f %s %.3f ms/step | e2e %.1f %s" % (d["value"], d["unit"], d["ms_per_step"], e.get("value", 0), e.get("ms_per_step")))
        if "roofline" in d: print("   ", {k: round(v, 3) for k, v in d["roofline"].get("kernel_ms_all", {}).items()}, d.get("cpu_baseline", {}).get("value"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
ls -la gpurun_out/*${TAG}*.ncu-rep
