#!/usr/bin/env python
"""Makes tests/golden/n3_ref.npz: read staging pinned by the REFERENCE'S OWN bamReadBuffer.

Inputs  = every alignment record of the reference's test BAM (test/S55_test_realigned.bam, 2115 records, kept as raw
          record fields: 4-bit bases, u32 CIGAR words) plus 600 synthetic records that reach the branches the BAM does
          not (secondary / unmapped / duplicate flags, low-quality reads, reverse reads with read-through, soft clips at
          both ends, single-end duplicates, records without sequence or qualities).
Outputs = what the reference's addReadToBuffer -> checkAndTrimRead (src/cython/cwindow.pyx:560-595, 332-481) and
          ReadArray.setWindowPointers (cwindow.pyx:208-236) make of them, run here through oracle/_ref
          (l3_ref_wrap.stage_reads), for the default options and three other option sets.

The record -> cAlignedRead field derivation of ReadIterator.get (htslibWrapper.pyx:328-406: nibble -> letter, pos minus a
leading soft clip, bam_endpos) needs htslib and cannot be executed; it is restated in `to_reference_tuple` below.

Run in the build container (needs /root/reference and oracle/_ref):  python tests/golden/make_n3_fixture.py
"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import oracle as O  # noqa: E402
from platypus_b200 import reads as R  # noqa: E402

BAM = "/root/reference/test/S55_test_realigned.bam"
NIB = b"=ACMGRSVTWYHKDBN"
OPTION_SETS = [
    {},
    {"trimReadFlank": 4, "minMapQual": 30, "minBaseQual": 25, "minGoodQualBases": 60},
    {"filterDuplicates": 0, "filterReadsWithUnmappedMates": 0, "filterReadsWithDistantMates": 0,
     "filterReadPairsWithSmallInserts": 0},
    {"trimOverlapping": 0, "trimAdapter": 0, "trimSoftClipped": 0},
]


def synthetic_records(n, seed, pos0):
    rng = random.Random(seed)
    core, cig, nib, qual = [], [], [], []
    pos = pos0
    prev = None
    for i in range(n):
        L = rng.choice([0, 30, 50, 76, 100, 101, 150])
        u = rng.random()
        flag = 0
        if rng.random() < 0.8:
            flag |= 0x1
            if rng.random() < 0.85:
                flag |= 0x2
            if rng.random() < 0.05:
                flag |= 0x8
            if rng.random() < 0.5:
                flag |= 0x20
        if rng.random() < 0.5:
            flag |= 0x10
        for bit, p in ((0x4, 0.03), (0x100, 0.03), (0x400, 0.05), (0x200, 0.02)):
            if rng.random() < p:
                flag |= bit
        mapq = rng.choice([0, 5, 19, 20, 29, 30, 60, 60, 60])
        pos += rng.choice([0, 0, 1, 3, 17])
        ops = []
        left = L
        if L and rng.random() < 0.3:
            k = rng.randint(1, min(12, L - 1))
            ops.append((4, k))
            left -= k
        tail = 0
        if left > 2 and rng.random() < 0.3:
            tail = rng.randint(1, min(12, left - 1))
            left -= tail
        if left > 10 and rng.random() < 0.3:
            a = rng.randint(1, left - 5)
            ops += [(0, a), (rng.choice([1, 2]), rng.randint(1, 3))]
            if ops[-1][0] == 1:
                left -= ops[-1][1]
            ops.append((0, max(1, left - a)))
        elif left > 0:
            ops.append((0, left))
        if tail:
            ops.append((4, tail))
        if L and rng.random() < 0.05:
            ops.append((5, 7))
        mate_ref = 0 if rng.random() < 0.95 else 1
        mate_pos = pos + rng.randint(-300, 300)
        tlen = rng.choice([0, L // 2, -(L // 2), L - 1, L + 40, -(L + 40), 2 * L - 3, 300, -300, rng.randint(-500, 500)])
        if prev is not None and u < 0.12:   # a copy of the previous record's position / length (duplicate rule)
            pos, L2, mp = prev
            if L2 == L:
                mate_pos = mp if rng.random() < 0.6 else mp + 1
        q = bytes(rng.choice([0, 2, 4, 5, 12, 19, 20, 30, 37, 41]) for _ in range(L))
        if L and rng.random() < 0.15:
            q = bytes(rng.choice([2, 3, 19]) for _ in range(L))      # mostly low: the LOW_QUAL_BASES filter
        if L and rng.random() < 0.03:
            q = b"\xff" * L                                            # no qualities stored
        codes = [rng.choice([1, 2, 4, 8, 1, 2, 4, 8, 15, 3]) if rng.random() < 0.03 else rng.choice([1, 2, 4, 8]) for _ in range(L)]
        packed = bytearray((L + 1) // 2)
        for k, c in enumerate(codes):
            packed[k >> 1] |= c << (4 * (1 - (k & 1)))
        core.append((0, pos, mapq, flag, mate_ref, mate_pos, tlen))
        cig.append(ops)
        nib.append(bytes(packed))
        qual.append(q)
        prev = (pos, L, mate_pos)
    return core, cig, nib, qual


def merge(rec, extra):
    core, cig, nib, qual = extra
    n0 = rec.n
    c = np.array(core, np.int64).reshape(-1, 7)
    cig_words = np.array([(ln << 4) | op for ops in cig for op, ln in ops] + [0], np.uint32)
    cig_off = np.concatenate([rec.cigar_off, rec.cigar_off[-1] + np.cumsum([len(o) for o in cig])]).astype(np.int64)
    seq_off = np.concatenate([rec.seq_off, rec.seq_off[-1] + np.cumsum([len(q) for q in qual])]).astype(np.int64)
    nib_off = np.concatenate([rec.nib_off, rec.nib_off[-1] + np.cumsum([len(x) for x in nib])]).astype(np.int64)
    return R.BamRecords(
        rec.ref_names, np.concatenate([rec.ref_id, c[:, 0]]).astype(np.int32), np.concatenate([rec.pos, c[:, 1]]).astype(np.int32),
        np.concatenate([rec.mapq, c[:, 2]]).astype(np.uint8), np.concatenate([rec.flag, c[:, 3]]).astype(np.uint16),
        np.concatenate([rec.mate_ref_id, c[:, 4]]).astype(np.int32), np.concatenate([rec.mate_pos, c[:, 5]]).astype(np.int32),
        np.concatenate([rec.tlen, c[:, 6]]).astype(np.int32), cig_off,
        np.concatenate([rec.cigar[:int(rec.cigar_off[-1])], cig_words]).astype(np.uint32), seq_off, nib_off,
        np.concatenate([rec.nib[:int(rec.nib_off[-1])], np.frombuffer(b"".join(nib) + b"\0", np.uint8)]),
        np.concatenate([rec.qual[:int(rec.seq_off[-1])], np.frombuffer(b"".join(qual) + b"\0", np.uint8)])), n0


def to_reference_tuple(rec, i):
    """Record i as the reference's ReadIterator.get would hand it on (htslibWrapper.pyx:328-406), or None."""
    b0, b1 = int(rec.seq_off[i]), int(rec.seq_off[i + 1])
    L = b1 - b0
    if L == 0 or rec.qual[b0] == 0xFF:
        return None
    nb = rec.nib[int(rec.nib_off[i]):int(rec.nib_off[i + 1])]
    seq = bytes(NIB[(nb[k >> 1] >> (4 * (1 - (k & 1)))) & 15] for k in range(L))
    cg = [(int(w) & 15, int(w) >> 4) for w in rec.cigar[int(rec.cigar_off[i]):int(rec.cigar_off[i + 1])]]
    pos = int(rec.pos[i]) - (cg[0][1] if cg and cg[0][0] == 4 else 0)
    ref_len = sum(n for op, n in cg if op in (0, 2, 3, 7, 8))
    end = int(rec.pos[i]) + (ref_len if ref_len > 0 else 1)
    return (seq, rec.qual[b0:b1].tobytes(), cg, int(rec.ref_id[i]), pos, end, int(rec.mapq[i]), int(rec.flag[i]),
            int(rec.mate_ref_id[i]), int(rec.mate_pos[i]), int(rec.tlen[i]))


def main():
    W = O.ref_l3()
    assert W is not None and hasattr(W, "stage_reads"), "oracle/_ref is not built (python -c 'from oracle import build; build.build_all()')"
    rec, n_bam = merge(R.read_bam_records(BAM), synthetic_records(600, 20261017, 32640000))
    tuples = [to_reference_tuple(rec, i) for i in range(rec.n)]
    kept = np.array([t is not None for t in tuples], np.uint8)
    live = [t for t in tuples if t is not None]
    lo, hi = min(t[4] for t in live), max(t[5] for t in live)
    rng = random.Random(7)
    starts = sorted(rng.randint(lo - 200, hi + 50) for _ in range(400))
    wins = [(s, s + rng.choice([1, 20, 60, 150, 400, 1500])) for s in starts]
    out = {"n_bam": n_bam, "kept": kept, "windows": np.array(wins, np.int32), "n_option_sets": len(OPTION_SETS)}
    for k in ("ref_id", "pos", "mapq", "flag", "mate_ref_id", "mate_pos", "tlen", "cigar_off", "cigar", "seq_off", "nib_off", "nib", "qual"):
        out["rec_" + k] = getattr(rec, k)
    raw = rec.qual
    for s, ov in enumerate(OPTION_SETS):
        r = W.stage_reads(live, lo, hi, wins if s == 0 else [], ov)
        good = np.zeros(rec.n, np.uint8)
        flag = rec.flag.copy()
        good[kept == 1] = r["good"]
        flag[kept == 1] = np.array(r["flag"], np.uint16)
        zeroed = []
        for j, i in enumerate(np.nonzero(kept)[0]):
            b0 = int(rec.seq_off[i])
            q = np.frombuffer(r["qual"][j], np.uint8)
            d = np.nonzero(q != raw[b0:b0 + len(q)])[0]
            assert (q[d] == 0).all()
            zeroed.append(d + b0)
        out["o%d_good" % s], out["o%d_flag" % s] = good, flag
        out["o%d_zeroed" % s] = np.concatenate(zeroed).astype(np.int64)
        out["o%d_counts" % s] = np.array(r["counts"], np.int32)
        out["o%d_options" % s] = np.frombuffer(repr(sorted(ov.items())).encode(), np.uint8)
        if s == 0:
            out["slices"] = np.array(r["windows"], np.int32)
        print("option set %d: %d good, %d bad, counts %s, %d qualities zeroed" %
              (s, int(good.sum()), int(kept.sum() - good.sum()), r["counts"], len(out["o%d_zeroed" % s])))
    path = os.path.join(ROOT, "tests", "golden", "n3_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", rec.n, "records,", int(kept.sum()), "reads,", len(wins), "windows")


if __name__ == "__main__":
    main()
