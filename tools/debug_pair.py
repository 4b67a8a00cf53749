#!/usr/bin/env python
"""Debug helper: one adversarial window through the GPU and the oracle, pair by pair."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

from oracle import oracle as O
from platypus_b200.engine import Engine
from tests import cases

w, n = int(sys.argv[1]), int(sys.argv[2]) if len(sys.argv) > 2 else 2300
b = cases.adversarial_batch(20261017, n)
eng = Engine(0)
lo = max(0, w - int(os.environ.get("CTX", "0")))
sub = b.slice_windows(lo, w + 1)
ll, sc = eng.window_loglik(sub)
st = eng.last_stats()
ll0, sc0, _ = O.window_loglik(sub)
bad = np.nonzero(sc != sc0)[0]
print("windows", lo, w, "pairs", len(sc), "bad", len(bad), st)
H = int(sub.win_hap_off[-1] - sub.win_hap_off[-2])
T = int(sub.wi_slot_off[-1] - sub.wi_slot_off[-2])
off = int(sub.ll_offsets()[-2])
print("last window: gpu vs oracle (rows = haplotypes)")
g, o = sc[off:].reshape(H, T), sc0[off:].reshape(H, T)
for t in range(T):
    if (g[:, t] != o[:, t]).any():
        print("read", t, "gpu", g[:, t], "oracle", o[:, t])
