"""
Generates the golden fixtures in this directory.  Run in the BUILD container, where
/root/reference exists; the GPU box only reads the committed .npz files.

  align_ref.npz   inputs + scores from the UNMODIFIED reference kernel
                  (src/c/align.c fastAlignmentRoutine via oracle/_ref/libalign_ref.so)
  calign_ref.npz  inputs + scores from the reference's src/cython/calign.pyx
                  mapAndAlignReadToHaplotype (oracle/_ref/calign*.so), gap-open tables from the
                  oracle's restatement of chaplotype.pyx:552-590
  window_restated.npz  a small multi-individual batch with per-read LL, GL, EM frequencies and
                  posteriors from the oracle (restatement of chaplotype/cgenotype/cpopulation:
                  "parity unpinned" above the integer score, see oracle/platypus_oracle.h)

usage: python tests/golden/make_golden.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402
from tests import cases  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def pack(list_of_bytes):
    off = np.zeros(len(list_of_bytes) + 1, np.int64)
    np.cumsum([len(b) for b in list_of_bytes], out=off[1:])
    return off, np.frombuffer(b"".join(list_of_bytes), np.uint8).copy()


def make_align(n=600, seed=11):
    assert O.ref_align_lib() is not None, "reference align.c not built (need /root/reference)"
    rng = random.Random(seed)
    haps, gos, reads, quals, scores = [], [], [], [], []
    for i in range(n):
        hap, go, read, qual = cases.random_alignment_case(rng, i)
        s_nt = O.ref_fast_align(hap, read, qual, go, traceback=False)
        s_tb = O.ref_fast_align(hap, read, qual, go, traceback=True)
        assert s_nt == s_tb, "traceback changed the score?"
        haps.append(hap[:len(read) + 15])
        gos.append(go[:len(read) + 15])
        reads.append(read)
        quals.append(qual)
        scores.append(s_nt)
    ho, hs = pack(haps)
    _, gs = pack(gos)
    ro, rs = pack(reads)
    _, qs = pack(quals)
    np.savez_compressed(os.path.join(HERE, "align_ref.npz"), hap_off=ho, hap=hs, gap_open=gs, read_off=ro, read=rs,
                        qual=qs, score=np.array(scores, np.int32))
    print("align_ref.npz:", n, "cases, score range", min(scores), max(scores))


def make_calign(n=500, seed=12):
    cw = O.ref_calign()
    assert cw is not None, "reference calign.pyx not built (need /root/reference + Cython)"
    rng = random.Random(seed)
    haps, reads, quals, rstart, hstart, scores = [], [], [], [], [], []
    for i in range(n):
        hap, read, qual, read_start, hap_start = cases.random_mapping_case(rng, i)
        go = O.gap_open(hap)
        s = cw.map_and_align(read, qual, read_start, hap_start, hap, go, 3, 2, 1, 0)
        haps.append(hap)
        reads.append(read)
        quals.append(qual)
        rstart.append(read_start)
        hstart.append(hap_start)
        scores.append(s)
    ho, hs = pack(haps)
    ro, rs = pack(reads)
    _, qs = pack(quals)
    np.savez_compressed(os.path.join(HERE, "calign_ref.npz"), hap_off=ho, hap=hs, read_off=ro, read=rs, qual=qs,
                        read_start=np.array(rstart, np.int32), hap_start=np.array(hstart, np.int32),
                        score=np.array(scores, np.int32))
    print("calign_ref.npz:", n, "cases;", sum(1 for s in scores if s == 1000000), "sentinel")


def make_window():
    batch = cases.edge_batch(seed=5)
    arrs, ll, sc, st = O.population_run(batch)
    out = {k: v for k, v in arrs.items() if k != "max_haps"}
    np.savez_compressed(os.path.join(HERE, "window_restated.npz"), ll=ll, score=sc, **out)
    print("window_restated.npz:", batch.n_windows, "windows", st)


if __name__ == "__main__":
    make_align()
    make_calign()
    make_window()
