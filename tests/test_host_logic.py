"""CPU tests of the host side: the packed-lane DP (host emulation of the s16x2 intrinsics), batch
packing / sharding, and that the C-ABI library loads and exports every symbol of
include/platypus_b200.h (no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from platypus_b200 import _abi, synth
from platypus_b200.batch import WindowBatch, shard_bounds
from tests import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_packed_lane_dp_matches_oracle_on_host(oracle, tmp_path):
    exe = str(tmp_path / "dp_lane_check")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "dp_lane_check.cpp"),
                           os.path.join(ROOT, "oracle", "libplatypus_oracle.so"),
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    out = subprocess.run([exe, "4000"], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout
    assert "mismatches 0 general 0" in out.stdout
    assert "flank mismatches 0 traceback mismatches 0" in out.stdout


def _lib():
    from __graft_entry__ import build
    build()
    from platypus_b200.engine import load_library
    return load_library()


def test_library_exports_every_declared_symbol():
    lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "platypus_b200.h")).read()
    declared = set(re.findall(r"\b(plb_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(_abi.EXPORTED_SYMBOLS), declared ^ set(_abi.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.plb_abi_version() == 6


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from platypus_b200.engine import Engine, PlbError
    with pytest.raises(PlbError) as e:
        Engine(0)
    assert "no CPU fallback" in str(e.value)


def test_ll_offsets_and_validate():
    lib = _lib()
    b = cases.edge_batch(seed=5)
    s = b.as_struct()
    off = np.zeros(b.n_windows * b.n_individuals + 1, np.int64)
    tot = C.c_int64()
    assert lib.plb_ll_offsets(C.byref(s), off.ctypes.data, C.byref(tot)) == 0
    assert np.array_equal(off, b.ll_offsets()) and tot.value == off[-1]
    opt = _abi.PlbOptions.default()
    assert lib.plb_validate(C.byref(s), C.byref(opt), 0) == 0
    assert lib.plb_validate(C.byref(s), C.byref(opt), 1) == _abi.PLB_ERR_SHAPE
    assert b"max_haps" in lib.plb_last_error()
    # both run-time modes are accepted (scope row a2 / N2); anything but 0/1 is an argument error
    for kw in (dict(calc_flank_score=1), dict(use_mapq_cap=1)):
        assert lib.plb_validate(C.byref(s), C.byref(_abi.PlbOptions.default(**kw)), 0) == 0
    opt2 = _abi.PlbOptions.default(calc_flank_score=2)
    assert lib.plb_validate(C.byref(s), C.byref(opt2), 0) == _abi.PLB_ERR_ARG
    # haplotype shorter than read + 15: the reference would read out of bounds (calign.pyx:256-259)
    from platypus_b200.batch import Read, Window
    w = Window(100, 140, 50, [b"ACGT" * 10], [([Read(b"A" * 30, bytes([30] * 30), 100, 130)], [], [])])
    b2 = WindowBatch.from_windows([w], 1)   # keep the batch alive: the struct points into its arrays
    s2 = b2.as_struct()
    assert lib.plb_validate(C.byref(s2), C.byref(opt), 0) == _abi.PLB_ERR_SHAPE


def test_batch_packing_shares_reads():
    b = cases.edge_batch(seed=5)
    assert b.n_slots > b.n_reads, "shared Read objects must be pooled once"
    assert b.wi_slot_off[-1] == b.n_slots == len(b.slot_read)
    assert (b.wi_n_good + b.wi_n_bad <= np.diff(b.wi_slot_off)).all()


def test_slice_equals_offset_generation():
    full = synth.make_batch(40, n_haps=4, n_reads=8, read_len=60, hap_len=120)
    part = synth.make_batch(13, n_haps=4, n_reads=8, read_len=60, hap_len=120, window_offset=20)
    sl = full.slice_windows(20, 33)
    for k in ("hap_seq", "read_seq", "read_qual", "read_pos", "read_end", "read_mapq", "win_start", "win_end",
              "hap_start", "hap_var_mask", "win_n_var"):
        assert np.array_equal(getattr(part, k), getattr(sl, k)), k
    assert shard_bounds(10, 4) == [0, 3, 6, 8, 10]


def test_synth_shapes_match_config2_accounting():
    b = synth.make_batch(3)
    assert b.n_haps == 24 and b.n_reads == 192
    assert synth.algorithmic_cells(b) == 3 * 8 * 64 * 150 * 16
    assert synth.algorithmic_bytes(b) == 3 * 17496
    v = synth.make_batch(5, read_len_range=(100, 250), hap_len_range=(200, 500))
    lens = np.diff(v.read_seq_off)
    assert lens.min() >= 100 and lens.max() <= 250
    hl = np.diff(v.hap_seq_off)
    assert (hl >= lens.max() + 16).all() and hl.max() <= 500


@pytest.mark.parametrize("prologue", [True, False])
def test_n1_selection_bookkeeping_replay(oracle, golden_dir, prologue, monkeypatch):
    """plb_select_replay_host (the library's trial sets / isHaplotypeValid / heap and sort replay / final ranking, no
    GPU) fed with the reference's own trial scores reproduces the reference's selection on the golden windows: every
    round asks for exactly the trial sets the reference scores, in its order, and the returned haplotypes match.  With the
    prologue (default) the score-independent trial sets of the first rounds are requested in one go."""
    from platypus_b200.engine import Engine
    lib = _lib()
    monkeypatch.setenv("PLB_SELECT_CHECK", "1")
    if not prologue:   # plain schedule: one scoring request per round, in the reference's order
        monkeypatch.setenv("PLB_SELECT_NO_PROLOGUE", "1")
    gold = cases.n1_golden_cases(golden_dir)
    by_opts = {}
    for g in gold:
        by_opts.setdefault(tuple(sorted(g["opts"].items())), []).append(g)
    n_windows = 0
    for key, group in by_opts.items():
        o = dict(key)
        cs = [cases.n1_window_case(g["seed"], g["drop"]) for g in group]
        batch, vset = cases.n1_batch(cs, [g["ref_seq"] for g in group], [g["hap_start"] for g in group])
        table = [dict(zip(g["trial_mask"], g["trial_score"])) for g in group]
        asked = [[] for _ in group]

        def score(hap_win, hap_mask):
            for w, m in zip(hap_win, hap_mask):
                asked[int(w)].append(int(m))
            return [table[int(w)][int(m)] for w, m in zip(hap_win, hap_mask)]
        sel = _abi.PlbSelectOptions(o["max_haplotypes"], o["original_max_haplotypes"], o["max_variants"], o["filter_by_coverage"],
                                    o["coverage_sampling_level"])
        out = Engine.select_replay(batch, vset, score, sel, lib=lib)
        for k, g in enumerate(group):
            n = int(out["n_sel"][k])
            assert [int(m) for m in out["sel_mask"][k, :n]] == g["sel_mask"], g["seed"]
            if prologue:   # the first rounds' trial sets are requested together (they do not depend on the scores)
                assert sorted(asked[k]) == sorted(g["trial_mask"]), g["seed"]
            else:
                assert asked[k] == g["trial_mask"], g["seed"]
            assert int(out["n_scored"][k]) == len(g["trial_mask"])
            n_windows += 1
    assert n_windows == len(gold)


def test_n1_selection_rejects_bad_variant_lists(golden_dir):
    from platypus_b200.engine import Engine, PlbError
    from platypus_b200.batch import VariantSet
    lib = _lib()
    g = cases.n1_golden_cases(golden_dir)[0]
    c = cases.n1_window_case(g["seed"], g["drop"])
    batch, _ = cases.n1_batch([c], [g["ref_seq"]], [g["hap_start"]])
    vs = c["variants"]
    for bad in ([vs[1], vs[0]] + vs[2:], vs + [vs[-1]], [(c["win_end"] + 5, b"A", b"C", 1)]):
        with pytest.raises(PlbError):
            Engine.select_replay(batch, VariantSet.from_lists([bad]), lambda w, m: [0.0] * len(w), lib=lib)


def test_n1_two_group_pipeline_matches_single_group(monkeypatch):
    """Large batches run as two groups of windows whose rounds alternate (host bookkeeping of one overlaps the scoring
    of the other); the result must not depend on the grouping.  Replay with a deterministic, tie-rich score."""
    from platypus_b200.engine import Engine
    lib = _lib()
    b1, v1 = synth.make_select_batch(40, n_vars=8, n_reads=8)
    b2, v2 = synth.make_select_batch(30, n_vars=6, n_reads=8, window_offset=40)

    def score(hw, hm):
        return -((hm * np.uint64(2654435761) + hw.astype(np.uint64) * np.uint64(97)) % np.uint64(23)).astype(np.float64)
    sel = _abi.PlbSelectOptions.default(max_haplotypes=12, original_max_haplotypes=14)
    outs = []
    for b, v in ((b1, v1), (b2, v2)):
        res = []
        for groups in ("1", "2"):
            monkeypatch.setenv("PLB_SELECT_GROUPS", groups)
            monkeypatch.setenv("PLB_SELECT_CHECK", "1")
            res.append(Engine.select_replay(b, v, score, sel, lib=lib))
        for k in ("n_sel", "sel_mask", "n_scored"):
            assert np.array_equal(res[0][k], res[1][k]), k
        assert np.array_equal(np.nan_to_num(res[0]["sel_score"]), np.nan_to_num(res[1]["sel_score"]))
        assert np.all(res[0]["n_sel"] == 11)
        outs.append(res[0])


def test_n1_haplotype_lengths_on_host(golden_dir):
    """plb_build_haplotypes_host with hap_seq = NULL only sizes the haplotypes (host walk of getMutatedSequence, no GPU):
    the offsets must be the lengths of the reference's own Haplotype.cHaplotypeSequence (tests/golden/n1_ref.npz)."""
    lib = _lib()
    gold = cases.n1_golden_cases(golden_dir)
    cs = [cases.n1_window_case(g["seed"], g["drop"]) for g in gold]
    batch, vset = cases.n1_batch(cs, [g["ref_seq"] for g in gold], [g["hap_start"] for g in gold])
    hap_win, hap_mask, want = [], [], []
    for k, g in enumerate(gold):
        hap_win.append(k)
        hap_mask.append(0)
        want.append(len(g["ref_seq"]))
        for m, seq in zip(g["sel_mask"], g["hap_seqs"]):
            hap_win.append(k)
            hap_mask.append(m)
            want.append(len(seq))
    hw, hm = np.asarray(hap_win, np.int32), np.asarray(hap_mask, np.uint64)
    off = np.zeros(len(hw) + 1, np.int64)
    s_, v_ = batch.as_struct(), vset.as_struct()
    rc = lib.plb_build_haplotypes_host(None, C.byref(s_), C.byref(v_), len(hw), _abi.ptr(hw), _abi.ptr(hm), _abi.ptr(off), None, 0)
    assert rc == 0, lib.plb_last_error()
    assert np.array_equal(np.diff(off), want)
    assert len(want) > 1000


def test_n1_bookkeeping_stress_with_tied_scores(oracle):
    """The library's heap / sort / tuple-order replay against CPython's own heapq and sorted (oracle/select_oracle.py) on
    random windows built to provoke the awkward cases: scores drawn from a handful of values (ties everywhere), several
    alleles at one position (unequal variants of which neither is 'less'), SNP + indel at one base, overlapping deletions
    (invalid combinations), heap capacities from 2 to 63 with maxHaplotypes above and below originalMaxHaplotypes."""
    import random
    from oracle import select_oracle as S
    from platypus_b200.batch import VariantSet, Window, WindowBatch
    from platypus_b200.engine import Engine
    lib = _lib()
    rng = random.Random(11)
    n_checked = 0
    for trial in range(250):
        ws = 1000
        we = ws + 40
        variants, pos = [], ws + 1
        while len(variants) < rng.randint(4, 10) and pos < we - 6:
            kind = rng.random()
            if kind < 0.45:
                for alt in sorted(rng.sample([b"A", b"C", b"G", b"T"], rng.choice([1, 2, 3]))):
                    variants.append((pos, b"N", alt))
            elif kind < 0.65:
                variants.append((pos, b"", b"ACG"[:rng.randint(1, 3)]))
            elif kind < 0.85:
                variants.append((pos, b"N" * rng.randint(1, 5), b""))
            else:
                variants.append((pos, b"NN", b"AC"))
            pos += rng.choice([0, 0, 1, 2, 3, 6])

        def vtype(v):
            nr, na = len(v[1]), len(v[2])
            return (0 if na == 1 else 1) if nr == na else 2 if nr == 0 else 3 if na == 0 else 4
        uniq = sorted(set(variants), key=lambda v: (v[0], vtype(v), len(v[1])))
        # equal keys keep a fixed order; drop exact duplicates of (pos, nRemoved, added)
        seen, vs = set(), []
        for v in uniq:
            k = (v[0], len(v[1]), v[2])
            if k not in seen:
                seen.add(k)
                vs.append((v[0], v[1], v[2], rng.choice([1, 1, 2, 3, 5])))
        vs = vs[:12]
        orig = rng.choice([3, 4, 6, 9, 17, 33, 50, 64])
        mx = rng.choice([orig, orig, max(2, orig - rng.randint(1, 3)), orig + rng.randint(1, 4)])
        levels = rng.choice([1, 2, 3, 7])
        salt = rng.randrange(1 << 30)

        def score_of_mask(m):
            return -float((m * 2654435761 + salt) % (1 << 32) % levels)
        ref = bytes(rng.choice(b"ACGT") for _ in range(240))
        w = S.SelectWindow(ref, ws, we, ws - 100, vs, [[]])
        want = S.select_haplotypes(w, mx, orig, 8, 0, 30, score_fn=lambda t: score_of_mask(sum(1 << v.idx for v in t)))
        batch = WindowBatch.from_windows([Window(ws, we, ws - 100, [ref], [([], [], [])])], 1)
        sel = _abi.PlbSelectOptions(mx, orig, 8, 0, 30)
        got = Engine.select_replay(batch, VariantSet.from_lists([vs]), lambda hw, hm: [score_of_mask(int(m)) for m in hm], sel,
                                   max_sel=4096, lib=lib)
        n = int(got["n_sel"][0])
        assert [int(m) for m in got["sel_mask"][0, :n]] == cases.masks_of([s_ for s_, _ in want]), (trial, orig, mx, levels)
        n_checked += n
    assert n_checked > 1500


def test_with_haplotypes_builds_a_valid_window_model_batch(oracle):
    """batch.with_haplotypes (the window-model batch after selection: reads of the reference batch, [reference] + selected
    haplotypes, masks as hap_var_mask) passes the library's host-side validation and runs through the oracle's window model."""
    from oracle import select_oracle as S
    from platypus_b200.batch import with_haplotypes
    lib = _lib()
    ref_batch, vset = synth.make_select_batch(5, n_vars=6, n_reads=12, read_len=80, n_individuals=2, seed=99)
    hap_off, seqs, masks = [0], [], []
    for w in range(ref_batch.n_windows):
        sw = S.window_from_batch(ref_batch, vset, w)
        for s_ in [()] + [x for x, _ in S.select_haplotypes(sw, 6, 6, 8, 1, 30)]:
            seqs.append(S.build_haplotype(sw.ref_seq, sw.win_start, sw.win_end, sw.hap_start, tuple(sw.vars[i] for i in s_)))
            masks.append(sum(1 << i for i in s_))
        hap_off.append(len(seqs))
    b = with_haplotypes(ref_batch, hap_off, seqs, masks, vset)
    assert b.n_haps == 30 and b.max_haps() == 6 and b.max_variants == 6
    assert b.read_seq is ref_batch.read_seq and b.slot_read is ref_batch.slot_read      # reads are shared, not copied
    s_ = b.as_struct()
    opt = _abi.PlbOptions.default()
    assert lib.plb_validate(C.byref(s_), C.byref(opt), 6) == 0, lib.plb_last_error()
    want, _, _, _ = oracle.population_run(b, max_haps=6)
    assert want["gl"].shape == (5, 2, 21) and np.all(np.isfinite(want["freq"]))
    np.testing.assert_allclose(want["freq"].sum(axis=1), 1.0, rtol=1e-12)
    # the sharded view of the variants used by run_select_sharded
    part = vset.slice_windows(2, 5)
    assert part.win_var_off[0] == 0 and part.n_vars(0) == vset.n_vars(2)
    assert np.array_equal(part.var_pos, vset.var_pos[vset.win_var_off[2]:vset.win_var_off[5]])


def test_pack_bases_and_nibbles_match_numpy():
    """plb_pack_bases_host / plb_pack_nibbles_host (staging row N3): 2-bit codes at any base offset, exceptions in order;
    BAM nibbles (htslibWrapper.pyx:414-416) give the same bytes as packing the decoded ASCII."""
    lib = _lib()
    rng = np.random.default_rng(11)
    nib_tab = np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    for n, base in ((0, 0), (1, 3), (37, 0), (150, 6), (1001, 2), (3 << 20, 1)):
        nib = rng.choice(np.array([1, 2, 4, 8, 15, 3, 0], np.uint8), size=n, p=[.24, .24, .24, .24, .02, .01, .01])
        ascii_ = nib_tab[nib]
        bam = np.zeros((n + 1) // 2 + 1, np.uint8)
        bam[:(n + 1) // 2] = (np.pad(nib, (0, n % 2))[0::2] << 4) | np.pad(nib, (0, n % 2))[1::2]
        want = np.zeros((base + n + 3) // 4 + 1, np.uint8)
        for jj in range(base):                             # earlier bases (all 'C') already packed in front
            want[jj >> 2] |= 1 << (2 * (jj & 3))
        pre = want.copy()
        exc = [(base + i, int(ch)) for i, ch in enumerate(ascii_) if int(ch) not in code]
        cd = np.array([code.get(int(ch), 0) for ch in ascii_], np.uint8) if n < 5000 else \
            np.select([ascii_ == 65, ascii_ == 67, ascii_ == 71, ascii_ == 84], [0, 1, 2, 3], 0).astype(np.uint8)
        j = base + np.arange(n)
        np.bitwise_or.at(want, j >> 2, (cd << (2 * (j & 3))).astype(np.uint8))
        for fn, src in ((lib.plb_pack_bases_host, ascii_), (lib.plb_pack_nibbles_host, bam)):
            dst = pre.copy()
            cap = len(exc) + 2
            pos, chr_ = np.full(cap, -1, np.int64), np.zeros(cap, np.uint8)
            k = C.c_int64(1)                                 # one entry already there: appended after it
            src = np.ascontiguousarray(src)
            assert fn(src.ctypes.data, n, dst.ctypes.data, base, pos.ctypes.data, chr_.ctypes.data, cap, C.byref(k)) == 0
            assert np.array_equal(dst, want), (n, base)
            assert k.value == 1 + len(exc)
            assert [(int(a), int(b)) for a, b in zip(pos[1:k.value], chr_[1:k.value])] == exc
            if exc:   # too small an exception array is an error, not an overrun
                k = C.c_int64(0)
                assert fn(src.ctypes.data, n, np.zeros_like(want).ctypes.data, base, pos.ctypes.data, chr_.ctypes.data,
                          len(exc) - 1, C.byref(k)) == _abi.PLB_ERR_SHAPE


def test_packed_batch_struct_and_validate():
    lib = _lib()
    b = cases.edge_batch(seed=5)
    p = b.pack(lib)
    assert p.seq_format == _abi.PLB_SEQ_2BIT and len(p.read_seq) == (int(b.read_seq_off[-1]) + 3) // 4 + 1
    assert len(p.read_exc_pos) == int(np.count_nonzero(~np.isin(b.read_seq[:int(b.read_seq_off[-1])], np.frombuffer(b"ACGT", np.uint8))))
    # unpacking on the host gives the original bytes back
    for pk, exc_p, exc_c, orig, n in ((p.read_seq, p.read_exc_pos, p.read_exc_chr, b.read_seq, int(b.read_seq_off[-1])),
                                      (p.hap_seq, p.hap_exc_pos, p.hap_exc_chr, b.hap_seq, int(b.hap_seq_off[-1]))):
        i = np.arange(n)
        back = np.frombuffer(b"ACGT", np.uint8)[(pk[i >> 2] >> (2 * (i & 3))) & 3].copy()
        back[exc_p] = exc_c
        assert np.array_equal(back, orig[:n])
    s = p.as_struct()
    assert lib.plb_validate(C.byref(s), C.byref(_abi.PlbOptions.default()), 0) == 0
    assert p.input_nbytes() < b.input_nbytes()
    bad = p.as_struct()
    bad.seq_format = 7
    assert lib.plb_validate(C.byref(bad), None, 0) == _abi.PLB_ERR_ARG


def test_validate_skips_unscored_reads():
    """A QC-fail read (or one that misses the window) is never aligned (chaplotype.pyx:343-361), so its length cannot
    make a batch invalid; the same read as a broken mate is scored and does."""
    from platypus_b200.batch import Read, Window
    lib = _lib()
    opt = _abi.PlbOptions.default()
    hap = b"ACGT" * 10
    long_qc = Read(b"A" * 30, bytes([30] * 30), 100, 130, 60, True)
    far = Read(b"A" * 30, bytes([30] * 30), 400, 430, 60, False)
    ok = WindowBatch.from_windows([Window(100, 140, 50, [hap], [([long_qc, far], [], [])])], 1)
    assert lib.plb_validate(C.byref(ok.as_struct()), C.byref(opt), 0) == 0
    bad = WindowBatch.from_windows([Window(100, 140, 50, [hap], [([], [], [far])])], 1)
    assert lib.plb_validate(C.byref(bad.as_struct()), C.byref(opt), 0) == _abi.PLB_ERR_SHAPE


def test_pack_quals_roundtrip_and_limits():
    """plb_pack_quals_host: 4 bits for <= 16 distinct qualities, 6 bits for <= 64, refused beyond (the batch stays at 8 bits);
    codes decode back to the bytes; WindowBatch.pack() carries table and width into the struct."""
    lib = _lib()
    rng = np.random.default_rng(3)
    for n, vals, want_bits in ((1, [30], 4), (1003, [2, 12, 23, 37], 4), (5 << 18, list(range(2, 41)), 6), (777, list(range(0, 64)), 6)):
        src = np.array(vals, np.uint8)[rng.integers(0, len(vals), n)]
        src[:len(vals)] = vals[:n]
        dst = np.zeros((n * 6 + 7) // 8 + 8, np.uint8)
        tab, bits = np.zeros(64, np.uint8), C.c_int32(0)
        assert lib.plb_pack_quals_host(src.ctypes.data, n, dst.ctypes.data, C.byref(bits), tab.ctypes.data) == 0
        assert bits.value == want_bits and list(tab[:len(set(src.tolist()))]) == sorted(set(src.tolist()))
        i = np.arange(n, dtype=np.int64)
        bit = i * bits.value
        word = dst[bit >> 3].astype(np.uint32) | (dst[(bit >> 3) + 1].astype(np.uint32) << 8)
        back = tab[(word >> (bit & 7).astype(np.uint32)) & ((1 << bits.value) - 1)]
        assert np.array_equal(back, src)
    too_many = np.arange(70, dtype=np.uint8)
    assert lib.plb_pack_quals_host(too_many.ctypes.data, 70, np.zeros(80, np.uint8).ctypes.data, C.byref(C.c_int32()),
                                   np.zeros(64, np.uint8).ctypes.data) == _abi.PLB_ERR_SHAPE
    bad = np.array([10, 94], np.uint8)
    assert lib.plb_pack_quals_host(bad.ctypes.data, 2, np.zeros(16, np.uint8).ctypes.data, C.byref(C.c_int32()),
                                   np.zeros(64, np.uint8).ctypes.data) == _abi.PLB_ERR_SHAPE
    b = synth.make_batch(3)
    p = b.pack(lib)
    assert p.qual_bits == 6 and p.seq_format == _abi.PLB_SEQ_2BIT and p.input_nbytes() < 0.6 * b.input_nbytes()
    s = p.as_struct()
    assert s.qual_bits == 6 and list(s.qual_table[:3]) == list(p.qual_table[:3])
    assert lib.plb_validate(C.byref(s), None, 0) == 0
    s.qual_bits = 5
    assert lib.plb_validate(C.byref(s), None, 0) == _abi.PLB_ERR_ARG
    assert b.pack(lib, quals=False).qual_bits == 0
