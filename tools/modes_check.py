"""Throughput of the two run-time modes (every alignment on the scalar path) on a GPU box.
usage: python tools/modes_check.py [windows]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from platypus_b200 import _abi, synth  # noqa: E402
from platypus_b200.engine import Engine  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
b = synth.make_batch(W)
eng = Engine(0)
for name, kw in (("default", {}), ("flank", dict(calc_flank_score=1)), ("hla", dict(use_mapq_cap=1))):
    opt = _abi.PlbOptions.default(**kw)
    eng.population_run(b, opt=opt)
    t = time.perf_counter()
    for _ in range(3):
        eng.population_run(b, opt=opt)
    dt = (time.perf_counter() - t) / 3
    st = eng.last_stats()
    print("%-8s %d windows: %.1f ms per call (pageable host buffers), %.0f GCUPS, %d alignments" %
          (name, W, dt * 1e3, st["cells"] / dt / 1e9, st["n_dp"]))
