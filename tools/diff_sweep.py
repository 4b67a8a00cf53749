#!/usr/bin/env python
"""Differential sweep beyond the fixed cases of the test suite: adversarial windows (tests/cases.adversarial_batch) for
the seeds given, GPU integer scores against the CPU oracle with the reference's own align.c (test infrastructure: this
tool, like the tests, is the only kind of code that touches oracle/).  usage: python tools/diff_sweep.py seed[:windows] ..."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from oracle import oracle as O
from platypus_b200.engine import Engine
from tests import cases

eng = Engine(0)
n_thr = os.cpu_count() or 1
worst = 0
for arg in sys.argv[1:] or ["1"]:
    seed, _, nw = arg.partition(":")
    b = cases.adversarial_batch(int(seed), int(nw or 1500))
    t0 = time.time()
    ll, sc = eng.window_loglik(b)
    packed = b.pack()
    ll2, sc2 = eng.window_loglik(packed)
    kind = O.use_reference_kernel(True, traceback=False)
    try:
        ll0, sc0, st0 = O.window_loglik(b, n_threads=n_thr)
    finally:
        O.use_reference_kernel(False)
    defined = sc0 < 15871
    bad = np.nonzero((sc != sc0) & defined)[0]
    bad2 = np.nonzero(sc2 != sc)[0]
    worst = max(worst, len(bad), len(bad2))
    print("seed %s: %d pairs, %d differ from the %s kernel, %d differ between packed and ASCII input, %.1f s" %
          (seed, len(sc), len(bad), "reference" if kind else "oracle", len(bad2), time.time() - t0), flush=True)
    if len(bad):
        print("   first:", bad[:8], sc[bad[:8]], sc0[bad[:8]])
sys.exit(1 if worst else 0)
