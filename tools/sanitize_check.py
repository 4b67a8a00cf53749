#!/usr/bin/env python
"""Small driver for compute-sanitizer: the pipelined host path with packed input (three chunks per queued job, tile lists in
HBM, unpack kernels, D2H stream) and the device-resident path, results compared with the one-call path.
usage: compute-sanitizer --tool memcheck python tools/sanitize_check.py [n_windows]"""
import sys

import numpy as np

sys.path.insert(0, ".")
from platypus_b200 import synth
from platypus_b200.engine import Engine
from tests import cases

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3200
eng = Engine(0)
batches = [synth.make_batch(n), cases.edge_batch(seed=4), synth.make_batch(n, read_len_range=(100, 250), hap_len_range=(200, 500))]
want = [eng.population_run(b) for b in batches]
packed = [b.pack() for b in batches]
jobs = []
got = []
for p in packed + packed:
    jobs.append(eng.population_submit(p))
    if len(jobs) == 2:
        got.append(eng.population_wait(jobs.pop(0)))
got.append(eng.population_wait(jobs.pop(0)))
for i, g in enumerate(got):
    w = want[i % len(want)]
    for k in ("gl", "freq", "em_post", "call", "var_phred"):
        assert np.array_equal(g[k], w[k]), (i, k)
print("sanitize_check ok:", len(got), "jobs,", n, "windows each")
