"""GPU parity tests: every check goes through the C ABI (platypus_b200.engine -> libplatypus_b200.so)
and compares with the CPU oracle / the committed golden vectors.  Integer scores must be
bit-exact; log-likelihoods, genotype likelihoods, frequencies and posteriors within 1e-4 relative
(the tolerance BASELINE.json states) - in practice they agree to ~1e-12."""
import os
import random

import numpy as np
import pytest

from platypus_b200 import _abi, synth
from platypus_b200.batch import Read, Window, WindowBatch
from tests import cases

pytestmark = pytest.mark.gpu
RTOL = 1e-4          # north_star tolerance
RTOL_TIGHT = 1e-9    # what identical summation order actually gives


def _unpack(off, data, i):
    return data[off[i]:off[i + 1]].tobytes()


def test_native_library_is_loaded(engine):
    assert os.path.basename(engine.lib._name) == "libplatypus_b200.so"
    assert engine.launch_count == 0


def test_s1_golden_align_ref(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "align_ref.npz"))
    n = len(g["score"])
    haps = [_unpack(g["hap_off"], g["hap"], i) for i in range(n)]
    gos = [_unpack(g["hap_off"], g["gap_open"], i) for i in range(n)]
    reads = [_unpack(g["read_off"], g["read"], i) for i in range(n)]
    quals = [_unpack(g["read_off"], g["qual"], i) for i in range(n)]
    got = engine.align_batch(haps, gos, reads, quals)
    assert np.array_equal(got, g["score"])
    assert engine.launch_count > 0


def test_s1_scalar_signature(engine, oracle):
    rng = random.Random(3)
    for i in range(6):
        hap, go, read, qual = cases.random_alignment_case(rng, i)
        assert engine.fast_align(hap, read, qual, go) == oracle.band_align(hap, read, qual, go)


def test_s1_fuzz_vs_oracle(engine, oracle):
    rng = random.Random(2024)
    cs = [cases.random_alignment_case(rng, i) for i in range(3000)]
    got = engine.align_batch([c[0][:len(c[2]) + 15] for c in cs], [c[1] for c in cs], [c[2] for c in cs],
                             [c[3] for c in cs])
    want = np.array([oracle.band_align(*[c[0], c[2], c[3], c[1]]) for c in cs], np.int32)
    assert np.array_equal(got, want)


def test_gap_open_vs_oracle(engine, oracle):
    rng = random.Random(5)
    haps = [cases.random_hap(rng, rng.randint(1, 700)) for _ in range(60)] + [b"A" * 120, b"ANNNA", b"G"]
    got = engine.gap_open(haps)
    for h, g in zip(haps, got):
        assert g == oracle.gap_open(h)


def _mapping_batch(cs):
    """One window per (hap, read) case so that calign golden cases run through S2."""
    wins = []
    for hap, read, qual, rs, hs in cs:
        r = Read(read, qual, rs, rs + len(read), 60)
        # window interval = whole haplotype so the overlap rule never fires; broken-mate list skips it anyway
        wins.append(Window(hs, hs + len(hap), hs, [hap], [([], [], [r])]))
    return WindowBatch.from_windows(wins, 1)


def test_s2_golden_calign_ref(engine, golden_dir):
    g = np.load(os.path.join(golden_dir, "calign_ref.npz"))
    n = len(g["score"])
    cs = []
    for i in range(n):
        cs.append((_unpack(g["hap_off"], g["hap"], i), _unpack(g["read_off"], g["read"], i),
                   _unpack(g["read_off"], g["qual"], i), int(g["read_start"][i]), int(g["hap_start"][i])))
    b = _mapping_batch(cs)
    ll, sc = engine.window_loglik(b)
    assert np.array_equal(sc, g["score"])


def test_s2_mapping_fuzz_vs_oracle(engine, oracle):
    rng = random.Random(77)
    cs = [cases.random_mapping_case(rng, i) for i in range(1500)]
    b = _mapping_batch(cs)
    ll, sc = engine.window_loglik(b)
    ll0, sc0, st0 = oracle.window_loglik(b)
    assert np.array_equal(sc, sc0)
    np.testing.assert_allclose(ll, ll0, rtol=RTOL_TIGHT, atol=0)


def _check_population(got, want, rtol=RTOL_TIGHT):
    for k in ("gl", "gl_log_max", "gof", "hap_like", "freq", "em_post"):
        np.testing.assert_allclose(got[k], want[k], rtol=rtol, atol=1e-300, err_msg=k)
    assert np.array_equal(got["call"], want["call"])
    assert np.array_equal(got["em_iters"], want["em_iters"])
    np.testing.assert_allclose(got["var_phred"], want["var_phred"], rtol=0, atol=0, err_msg="var_phred")


def test_s3_edge_batch_vs_oracle_and_golden(engine, oracle, golden_dir):
    b = cases.edge_batch(seed=5)
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, st0 = oracle.population_run(b)
    assert np.array_equal(got["score"], sc0)
    np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
    _check_population(got, want)
    g = np.load(os.path.join(golden_dir, "window_restated.npz"))
    assert np.array_equal(got["score"], g["score"])
    np.testing.assert_allclose(got["gl"], g["gl"], rtol=RTOL, atol=1e-300)
    st = engine.last_stats()
    assert st["n_pairs"] == st0["n_pairs"] and st["n_pairs_scored"] == st0["n_pairs_scored"]
    assert st["cells"] == st0["cells"]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_s3_more_edge_batches(engine, oracle, seed):
    b = cases.edge_batch(seed=seed, n_windows=20, n_individuals=4)
    got = engine.population_run(b, want_ll=True, max_haps=8)
    want, ll0, sc0, _ = oracle.population_run(b, max_haps=8)
    assert np.array_equal(got["score"], sc0)
    _check_population(got, want)


def test_s3_synth_config2_shape_sample(engine, oracle):
    """Config-2 shaped windows (8 haplotypes x 64 reads, 150 bp x 250 bp)."""
    b = synth.make_batch(96)
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, st0 = oracle.population_run(b, n_threads=os.cpu_count() or 1)
    assert np.array_equal(got["score"], sc0)
    np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
    _check_population(got, want)


def test_s3_synth_config3_ragged(engine, oracle):
    """Config-3 shaped windows: read length 100-250, haplotype length 200-500."""
    b = synth.make_batch(48, read_len_range=(100, 250), hap_len_range=(200, 500))
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, _ = oracle.population_run(b, n_threads=os.cpu_count() or 1)
    assert np.array_equal(got["score"], sc0)
    _check_population(got, want)


def test_s3_multi_individual(engine, oracle):
    b = synth.make_batch(16, n_haps=6, n_reads=10, n_individuals=12, read_len=100, hap_len=220)
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, _ = oracle.population_run(b)
    assert np.array_equal(got["score"], sc0)
    _check_population(got, want, rtol=1e-7)   # newFreq sums over individuals are reduced per thread


def test_s3_many_individuals(engine, oracle):
    """Multi-sample shape (config 5 in miniature): 300 individuals x 4 reads per window, so that a
    window spans several tiles and the EM / posterior kernels run their multi-individual path."""
    b = synth.make_batch(5, n_haps=5, n_reads=4, n_individuals=300, read_len=80, hap_len=200)
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, _ = oracle.population_run(b, n_threads=os.cpu_count() or 1)
    assert np.array_equal(got["score"], sc0)
    np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
    for k in ("gl", "gof", "hap_like"):
        np.testing.assert_allclose(got[k], want[k], rtol=RTOL_TIGHT, atol=1e-300, err_msg=k)
    # the sums over the 300 individuals run in the reference's order: integer outputs are equal, the rest to 1e-9
    np.testing.assert_allclose(got["freq"], want["freq"], rtol=RTOL_TIGHT, atol=0)
    np.testing.assert_allclose(got["em_post"], want["em_post"], rtol=RTOL_TIGHT, atol=1e-300)
    assert np.array_equal(got["call"], want["call"]) and np.array_equal(got["em_iters"], want["em_iters"])
    assert np.array_equal(got["var_phred"], want["var_phred"])


def test_s3_many_samples_golden_ref(engine, oracle, golden_dir):
    """k_population (thread per individual) at 300 and 2000 individuals - BASELINE config 5's sample count - against the
    reference's own Population class: the cross-individual sums run in the reference's order (cpopulation.pyx:436-447,
    546-581), so genotype calls, rounded variant posteriors AND the EM iteration count are equal, not merely close."""
    for b, want, use_em in cases.l3_pop_many_cases(golden_dir):
        opt = _abi.PlbOptions.default(use_em_likelihoods=use_em)
        got = engine.population_run(b, opt=opt)
        cases.check_l3_pop_many(got, want)
        ref, _, _, _ = oracle.population_run(b, opt)
        assert np.array_equal(got["em_iters"], ref["em_iters"]) and np.array_equal(got["call"], ref["call"])
        np.testing.assert_allclose(got["freq"], ref["freq"], rtol=1e-12, atol=0)


def test_long_haplotypes_and_reads(engine, oracle):
    """Realistic flanked haplotypes (~1 kb) and 250-400 bp reads: 16-bit vote counters,
    global-memory vote arrays, several haplotype groups per window."""
    rng = random.Random(9)
    wins = []
    for w in range(6):
        hl = rng.choice([900, 1400, 2100])
        ref = cases._rand_seq(rng, hl)
        haps = [ref]
        for _ in range(5):
            h = bytearray(ref)
            p = hl // 2 + rng.randint(-20, 20)
            h[p] = rng.choice([c for c in b"ACGT" if c != h[p]])
            haps.append(bytes(h))
        reads = []
        for k in range(20):
            L = rng.choice([250, 300, 400])
            idx = rng.randint(0, hl - L - 16)
            seq = cases.mutate(rng, haps[rng.randrange(6)][idx:], L)
            reads.append(Read(seq, bytes(rng.randint(2, 40) for _ in range(L)), 1000 + idx, 1000 + idx + L, 60))
        wins.append(Window(1000 + hl // 2 - 30, 1000 + hl // 2 + 30, 1000, haps, [(reads, [], [])]))
    b = WindowBatch.from_windows(wins, 1)
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, _ = oracle.population_run(b)
    assert np.array_equal(got["score"], sc0)
    _check_population(got, want)


def test_empty_and_degenerate_batches(engine, oracle):
    # windows where nobody has reads, and a zero-window batch
    w = Window(100, 140, 50, [b"ACGT" * 30, b"ACGA" * 30], [([], [], []), ([], [], [])])
    b = WindowBatch.from_windows([w, w], 2)
    got = engine.population_run(b)
    assert (got["gl"][:, :, :3] == 1.0).all() and (got["call"] == -1).all()
    empty = synth.make_batch(0)
    ll, sc = engine.window_loglik(empty)
    assert len(ll) == 0


def test_limits_are_errors_not_crashes(engine):
    from platypus_b200.engine import PlbError
    b = cases.edge_batch(seed=5)
    with pytest.raises(PlbError) as e:
        engine.population_run(b, max_haps=1)
    assert e.value.code == _abi.PLB_ERR_SHAPE
    with pytest.raises(PlbError) as e:
        engine.population_run(b, opt=_abi.PlbOptions.default(use_mapq_cap=2))
    assert e.value.code == _abi.PLB_ERR_ARG


def test_full_size_config2_properties(engine, oracle):
    """BASELINE config 2 at full size (10k windows x 8 haplotypes x 64 reads) through the C ABI.
    The oracle cannot score 5.12 M pairs in seconds, so check size-independent properties plus a
    random sample of windows against the oracle."""
    W = 10000
    b = synth.make_batch(W)
    got = engine.population_run(b, want_ll=True)
    st = engine.last_stats()
    assert st["n_pairs"] == W * 8 * 64 and st["cells"] == synth.algorithmic_cells(b)
    ll, sc = got["ll"], got["score"]
    assert (sc >= 0).all() and (sc < 15872).all() and (ll <= 0).all() and (ll >= -300).all()
    # rescaled genotype likelihoods peak at exactly 1, frequencies sum to 1, calls maximise GL
    np.testing.assert_array_equal(got["gl"].max(axis=2), 1.0)
    np.testing.assert_allclose(got["freq"].sum(axis=1), 1.0, rtol=1e-12)
    assert np.array_equal(got["call"][:, 0], got["gl"][:, 0, :].argmax(axis=1))
    # homozygous genotype log-likelihood = sum of that haplotype's read LLs (linearity check)
    L = ll.reshape(W, 8, 64)
    hom = np.array([0, 8, 15, 21, 26, 30, 33, 35])
    gl_log = np.log(got["gl"][:, 0, :]) + got["gl_log_max"]
    mask = got["gl"][:, 0, hom] > 1e-290
    np.testing.assert_allclose(gl_log[:, hom][mask], L.sum(axis=2)[mask], rtol=1e-9)
    # no cross-window state: reversing the window order reverses the outputs bit for bit
    idx = np.arange(W)[::-1]
    sub = [b.slice_windows(int(i), int(i) + 1) for i in idx[:300]]
    rev = _concat(sub)
    got_r = engine.population_run(rev, want_ll=True)
    assert np.array_equal(got_r["score"].reshape(300, -1), sc.reshape(W, -1)[idx[:300]])
    assert np.array_equal(got_r["gl"], got["gl"][idx[:300]])
    # random sample of windows against the oracle
    rng = np.random.default_rng(1)
    pick = np.sort(rng.choice(W, 40, replace=False))
    for w in pick:
        s = b.slice_windows(int(w), int(w) + 1)
        want, ll0, sc0, _ = oracle.population_run(s)
        assert np.array_equal(sc.reshape(W, -1)[w], sc0)
        np.testing.assert_allclose(got["gl"][w], want["gl"][0], rtol=RTOL_TIGHT, atol=1e-300)
        np.testing.assert_allclose(got["freq"][w], want["freq"][0], rtol=RTOL_TIGHT)
        np.testing.assert_array_equal(got["var_phred"][w, :want["var_phred"].shape[1]], want["var_phred"][0])


def _concat(batches):
    """Concatenate single-window batches (test helper)."""
    wins = []
    for s in batches:
        haps = [s.hap_seq[s.hap_seq_off[h]:s.hap_seq_off[h + 1]].tobytes() for h in range(s.n_haps)]
        reads = []
        for t in range(s.n_slots):
            r = int(s.slot_read[t])
            a, e = int(s.read_seq_off[r]), int(s.read_seq_off[r + 1])
            reads.append(Read(s.read_seq[a:e].tobytes(), s.read_qual[a:e].tobytes(), int(s.read_pos[r]),
                              int(s.read_end[r]), int(s.read_mapq[r]), bool(s.read_qcfail[r])))
        wins.append(Window(int(s.win_start[0]), int(s.win_end[0]), int(s.hap_start[0]), haps, [(reads, [], [])],
                           hap_var_mask=[int(m) for m in s.hap_var_mask],
                           var_prior=list(s.var_prior[0][:int(s.win_n_var[0])])))
    return WindowBatch.from_windows(wins, 1)


# ---- scope row a2 / N2: traceback, calculateFlankScore, HLA map-qual cap ------------------------

def _tb_cases(g, n):
    haps = [_unpack(g["hap_off"], g["hap"], i) for i in range(n)]
    gos = [_unpack(g["hap_off"], g["gap_open"], i) for i in range(n)]
    reads = [_unpack(g["read_off"], g["read"], i) for i in range(n)]
    quals = [_unpack(g["read_off"], g["qual"], i) for i in range(n)]
    return haps, gos, reads, quals


def test_s1_traceback_golden_ref(engine, golden_dir):
    """plb_fast_align / plb_align_traceback_host reproduce the reference's alignment rows, firstpos and
    score (src/c/align.c:523-577 run here at fixture time)."""
    g = np.load(os.path.join(golden_dir, "align_tb_ref.npz"))
    n = len(g["score"])
    haps, gos, reads, quals = _tb_cases(g, n)
    got = engine.align_traceback_batch([h[:len(r) + 15] for h, r in zip(haps, reads)], gos, reads, quals)
    for i in range(n):
        want = (int(g["score"][i]), _unpack(g["aln_off"], g["aln1"], i), _unpack(g["aln_off"], g["aln2"], i),
                int(g["firstpos"][i]))
        assert got[i] == want, "golden case %d" % i
    # the scalar signature, traceback requested through aln1/aln2 like the reference (align.c:96)
    for i in range(0, n, 57):
        assert engine.fast_align_traceback(haps[i], reads[i], quals[i], gos[i]) == got[i]


def test_s1_flank_score_golden_ref(engine, golden_dir):
    """One-pass flank score == calculateFlankScore of the reference's traceback (align.c:593-644)."""
    g = np.load(os.path.join(golden_dir, "align_tb_ref.npz"))
    n = len(g["score"])
    haps, gos, reads, quals = _tb_cases(g, n)
    sc, fl = engine.align_flank_batch(haps, gos, [0] * n, list(g["hap_flank"]), reads, quals)
    assert np.array_equal(sc, g["score"])
    assert np.array_equal(fl, g["flank_score"])


def test_s1_traceback_and_flank_fuzz_vs_oracle(engine, oracle):
    rng = random.Random(77)
    cs = [cases.random_alignment_case(rng, i) for i in range(1500)]
    got = engine.align_traceback_batch([c[0][:len(c[2]) + 15] for c in cs], [c[1] for c in cs], [c[2] for c in cs],
                                       [c[3] for c in cs])
    starts, flanks, want_f = [], [], []
    for i, c in enumerate(cs):
        assert got[i] == oracle.band_align_tb(c[0], c[2], c[3], c[1]), "case %d" % i
        st = rng.randint(0, len(c[0]) - len(c[2]) - 15)
        fk = rng.randint(1, max(1, len(c[0]) // 2))
        starts.append(st)
        flanks.append(fk)
        want_f.append(oracle.band_align_flank(c[0], c[2], c[3], c[1], st, fk))
    sc, fl = engine.align_flank_batch([c[0] for c in cs], [c[1] for c in cs], starts, flanks, [c[2] for c in cs],
                                      [c[3] for c in cs])
    assert [(int(a), int(b)) for a, b in zip(sc, fl)] == want_f


def _modes_mapping_batch(g, i0, i1, do_flank):
    """One window per golden case of calign_modes_ref.npz with the wanted flank setting; HLA-clipped
    cases cannot be expressed as windows (their clip is given, not derived) and are skipped."""
    wins, idx = [], []
    for i in range(i0, i1):
        if int(g["do_flank"][i]) != do_flank:
            continue
        read = _unpack(g["read_off"], g["read"], i)
        if read != _unpack(g["hash_read_off"], g["hash_read"], i) or len(read) < 7:
            continue
        hap = _unpack(g["hap_off"], g["hap"], i)
        hs, fl = int(g["hap_start"][i]), int(g["hap_flank"][i])
        rd = Read(read, _unpack(g["read_off"], g["qual"], i), int(g["read_start"][i]), int(g["read_start"][i]) + len(read), 60)
        # broken-mate list: scored without the overlap test (chaplotype.pyx:363-370)
        wins.append(Window(hs + fl, hs + fl + 1, hs, [hap], [([], [], [rd])]))
        idx.append(i)
    return WindowBatch.from_windows(wins, 1), idx


def test_s2_flank_mode_golden_calign_ref(engine, golden_dir):
    """Window path with calc_flank_score = 1 against the reference's mapAndAlignReadToHaplotype
    (doCalculateFlankScore = 1) outputs."""
    g = np.load(os.path.join(golden_dir, "calign_modes_ref.npz"))
    n = len(g["score"])
    b, idx = _modes_mapping_batch(g, 0, n, 1)
    assert len(idx) > 100
    ll, sc = engine.window_loglik(b, opt=_abi.PlbOptions.default(calc_flank_score=1))
    assert np.array_equal(sc, g["score"][idx])


@pytest.mark.parametrize("mode", [dict(calc_flank_score=1), dict(use_mapq_cap=1),
                                  dict(calc_flank_score=1, use_mapq_cap=1)])
def test_s3_modes_edge_batch_vs_oracle_and_golden(engine, oracle, golden_dir, mode):
    name = {(1, 0): "flank", (0, 1): "hla", (1, 1): "both"}[(mode.get("calc_flank_score", 0), mode.get("use_mapq_cap", 0))]
    opt = _abi.PlbOptions.default(**mode)
    b = cases.edge_batch(seed=5, overhang=True)
    got = engine.population_run(b, opt=opt, want_ll=True)
    want, ll0, sc0, st0 = oracle.population_run(b, opt)
    assert np.array_equal(got["score"], sc0)
    np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
    _check_population(got, want)
    g = np.load(os.path.join(golden_dir, "window_modes_restated.npz"))
    assert np.array_equal(got["score"], g[name + "_score"])
    np.testing.assert_allclose(got["gl"], g[name + "_gl"], rtol=RTOL, atol=1e-300)
    st = engine.last_stats()
    assert st["cells"] == st0["cells"] and st["n_pairs_scored"] == st0["n_pairs_scored"]
    for seed in (2, 3):
        b = cases.edge_batch(seed=seed, n_windows=20, n_individuals=4, overhang=True)
        got = engine.population_run(b, opt=opt, want_ll=True, max_haps=8)
        want, ll0, sc0, _ = oracle.population_run(b, opt, max_haps=8)
        assert np.array_equal(got["score"], sc0), seed
        _check_population(got, want)


def test_s3_modes_synth_sample(engine, oracle):
    """Config-2 shaped windows under both run-time modes (every alignment on the scalar path,
    queues sized for it), device-resident and host paths."""
    b = synth.make_batch(64)
    for mode in (dict(calc_flank_score=1), dict(use_mapq_cap=1)):
        opt = _abi.PlbOptions.default(**mode)
        got = engine.population_run(b, opt=opt, want_ll=True)
        want, ll0, sc0, _ = oracle.population_run(b, opt, n_threads=os.cpu_count() or 1)
        assert np.array_equal(got["score"], sc0), mode
        np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
        _check_population(got, want)
    # the default mode still takes the packed path afterwards (mode state does not leak)
    got = engine.population_run(b, want_ll=True)
    want, ll0, sc0, _ = oracle.population_run(b, n_threads=os.cpu_count() or 1)
    assert np.array_equal(got["score"], sc0)


def test_general_path_wavefront_vs_oracle(engine, oracle):
    """Windows whose haplotypes carry a byte outside ACGTN take the general path: every alignment is queued
    and runs as a 16-lane anti-diagonal wavefront (band_dp_wave16).  Config-2 shapes plus ragged lengths."""
    for b in (synth.make_batch(48), synth.make_batch(40, read_len_range=(100, 250), hap_len_range=(200, 500))):
        hs = b.hap_seq.copy()
        for w in range(b.n_windows):      # an IUPAC byte in the first haplotype of every window
            h = int(b.win_hap_off[w])
            hs[int(b.hap_seq_off[h]) + 5] = ord("R")
        b.hap_seq = hs
        got = engine.population_run(b, want_ll=True)
        want, ll0, sc0, _ = oracle.population_run(b, n_threads=os.cpu_count() or 1)
        assert np.array_equal(got["score"], sc0)
        np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
        _check_population(got, want)


# ---- scope row N4: per-site genotype calls -------------------------------------------------------

@pytest.mark.parametrize("n_ind", [1, 3, 30])
def test_n4_site_genotypes_vs_oracle(engine, oracle, n_ind):
    """plb_site_genotypes_host against the restatement of computeGenotypeCallAndLikelihoods; 30 individuals
    switch on the EM-frequency weighting (nIndividuals > 25, vcfutils.pyx:264-267)."""
    b = cases.edge_batch(seed=11, n_windows=16, n_individuals=n_ind)
    pop = engine.population_run(b)
    sites = cases.sites_for_batch(b, seed=n_ind)
    got = engine.site_genotypes(b, pop, sites)
    want = oracle.site_genotypes(b, pop, sites)
    for k in ("phased", "phred", "gt"):
        assert np.array_equal(got[k], want[k]), k
    for k in ("lik", "post", "gof", "gl_log10"):
        np.testing.assert_allclose(got[k], want[k], rtol=RTOL_TIGHT, atol=0, equal_nan=True, err_msg=k)


def test_n4_site_genotypes_synth(engine, oracle):
    b = synth.make_batch(200)
    pop = engine.population_run(b)
    sites = cases.sites_for_batch(b, seed=9)
    got = engine.site_genotypes(b, pop, sites)
    want = oracle.site_genotypes(b, pop, sites)
    for k in ("phased", "phred", "gt"):
        assert np.array_equal(got[k], want[k]), k
    np.testing.assert_allclose(got["lik"], want["lik"], rtol=RTOL_TIGHT, atol=0)
    called = (got["gt"][:, 0, 0] > 0) | (got["gt"][:, 0, 1] > 0)
    assert called.any() and not called.all()


def test_config1_hla_bam_windows_golden_ref(engine, oracle, golden_dir):
    """BASELINE config 1 on the GPU: real reads of the reference's test BAM against HLA-A allele haplotypes,
    integer scores equal to the reference's calign.pyx in all three modes, LL / GL equal to the oracle."""
    from tests.test_oracle import _check_hla_fixture
    b, g = cases.hla_fixture_batch(golden_dir)
    _check_hla_fixture(lambda kw: engine.window_loglik(b, opt=_abi.PlbOptions.default(**kw))[1], g)
    for kw in ({}, dict(use_mapq_cap=1), dict(calc_flank_score=1)):
        opt = _abi.PlbOptions.default(**kw)
        got = engine.population_run(b, opt=opt, want_ll=True)
        want, ll0, sc0, _ = oracle.population_run(b, opt)
        assert np.array_equal(got["score"], sc0)
        np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
        _check_population(got, want)


def test_l3_golden_ref(engine, golden_dir):
    """GPU per-read log-likelihoods, genotype likelihoods, GOF and hapLike against outputs of the reference's own
    chaplotype.pyx / cgenotype.pyx (tests/golden/l3_ref.npz), all four mode combinations."""
    n = 0
    for b, want in cases.l3_golden_cases(golden_dir):
        for (hla, flank), (w_ll, w_geno) in want.items():
            got = engine.population_run(b, opt=_abi.PlbOptions.default(use_mapq_cap=hla, calc_flank_score=flank), want_ll=True)
            # without the map-quality cap a log-likelihood is two IEEE operations on the integer score and a per-mapq
            # constant taken from the host's libm: bit-identical to the reference's double
            cases.check_l3(got["ll"], got, w_ll, w_geno, exact_ll=(hla == 0))
            n += w_ll.size
    assert n > 9000


def test_l3_population_golden_ref(engine, golden_dir):
    """GPU window model (rescale, EM, genotype calls, variant posteriors) against outputs of the reference's own
    Population class (tests/golden/l3_pop_ref.npz): 1-8 individuals, all mode combinations, both call rules."""
    n = 0
    for b, want, use_em, (hla, flank) in cases.l3_pop_golden_cases(golden_dir):
        opt = _abi.PlbOptions.default(use_mapq_cap=hla, calc_flank_score=flank, use_em_likelihoods=use_em)
        got = engine.population_run(b, opt=opt)
        cases.check_l3_pop(got, want, rtol=RTOL_TIGHT)
        n += 1
    assert n >= 40


def test_n4_golden_ref(engine, golden_dir):
    """k_site_genotypes against outputs of the reference's own computeGenotypeCallAndLikelihoods (n4_ref.npz)."""
    g = np.load(os.path.join(golden_dir, "n4_ref.npz"))
    for k, (b, sites) in enumerate(cases.n4_cases()):
        pop = engine.population_run(b)
        cases.check_n4(engine.site_genotypes(b, pop, sites), g, k, rtol=RTOL_TIGHT)


# ---- N1: haplotype construction + selection loop ------------------------------------------------------------------

def _n1_groups(golden_dir):
    by_opts = {}
    for g in cases.n1_golden_cases(golden_dir):
        # windows of one batch share the options and the number of individuals (no padding individuals: one without reads
        # scores 0.0 in computeBestScoreForHaplotype)
        n_ind = len(cases.n1_window_case(g["seed"], g["drop"])["per_ind"])
        by_opts.setdefault((n_ind,) + tuple(sorted(g["opts"].items())), []).append(g)
    out = []
    for key, group in by_opts.items():
        cs = [cases.n1_window_case(g["seed"], g["drop"]) for g in group]
        batch, vset = cases.n1_batch(cs, [g["ref_seq"] for g in group], [g["hap_start"] for g in group])
        o = dict(key[1:])
        sel = _abi.PlbSelectOptions(o["max_haplotypes"], o["original_max_haplotypes"], o["max_variants"], o["filter_by_coverage"],
                                    o["coverage_sampling_level"])
        out.append((group, batch, vset, sel))
    return out


def test_n1_build_haplotypes_golden_ref(engine, golden_dir):
    """k_build_haps vs Haplotype.cHaplotypeSequence of the reference (tests/golden/n1_ref.npz): every selected variant
    set of every golden window, the reference haplotype (mask 0) and every trial set of the rounds."""
    from oracle import select_oracle as S
    n = 0
    for group, batch, vset, _ in _n1_groups(golden_dir):
        hap_win, hap_mask, want = [], [], []
        for k, g in enumerate(group):
            c = cases.n1_window_case(g["seed"], g["drop"])
            w = cases.n1_select_window(c, g["ref_seq"], g["hap_start"])
            hap_win.append(k)
            hap_mask.append(0)
            want.append(g["ref_seq"])
            for m, seq in zip(g["sel_mask"], g["hap_seqs"]):
                hap_win.append(k)
                hap_mask.append(m)
                want.append(seq)
            for m in g["trial_mask"][:40]:     # trial sets: against the oracle's builder (itself pinned on the selected ones)
                hap_win.append(k)
                hap_mask.append(m)
                want.append(S.build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start,
                                              tuple(v for v in w.vars if m >> v.idx & 1)))
        got = engine.build_haplotypes(batch, vset, hap_win, hap_mask)
        assert got == want
        n += len(want)
    assert n > 2000


def test_n1_select_golden_ref(engine, golden_dir):
    """plb_select_haplotypes_host vs the reference's getFilteredHaplotypes on the golden windows: the same variant sets
    in the same order (the fixture is full of exactly tied scores), scores within 1e-9 of computeBestScoreForGenotype."""
    n_sel = 0
    for group, batch, vset, sel in _n1_groups(golden_dir):
        out = engine.select_haplotypes(batch, vset, sel)
        for k, g in enumerate(group):
            n = int(out["n_sel"][k])
            assert [int(m) for m in out["sel_mask"][k, :n]] == g["sel_mask"], g["seed"]
            assert int(out["n_scored"][k]) == len(g["trial_mask"])
            table = dict(zip(g["trial_mask"], g["trial_score"]))
            for j in range(n):
                if g["trial_mask"]:
                    np.testing.assert_allclose(out["sel_score"][k, j], table[int(out["sel_mask"][k, j])], rtol=RTOL_TIGHT)
                else:
                    assert np.isnan(out["sel_score"][k, j])
            n_sel += n
    assert n_sel > 500


def test_n1_best_score_haplotypes_golden_ref(engine, golden_dir):
    """plb_best_score_haplotypes_host vs the reference's computeBestScoreForHaplotype (variantFilter.pyx:212-234) on the
    reference haplotype and every selected haplotype of the golden windows (some individuals / windows have no reads)."""
    from platypus_b200.batch import with_haplotypes
    n = 0
    for group, batch, vset, _ in _n1_groups(golden_dir):
        hap_off, seqs, want = [0], [], []
        for g in group:
            seqs += [g["ref_seq"]] + g["hap_seqs"]
            want += [g["ref_hap_score"]] + list(g["hap_score"])
            hap_off.append(len(seqs))
        got = engine.best_score_haplotypes(with_haplotypes(batch, hap_off, seqs, None))
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=0)
        n += len(want)
    assert n > 1000


def test_n1_select_flank_mode_vs_oracle(engine, oracle):
    """options.calculateFlankScore reaches alignSingleRead inside the selection loop as well (chaplotype.pyx:384)."""
    from oracle import select_oracle as S
    batch, vset = synth.make_select_batch(12, n_vars=7, n_reads=24, read_len=100, hap_len=250, seed=77)
    opt = _abi.PlbOptions.default(calc_flank_score=1)
    sel = _abi.PlbSelectOptions.default(max_haplotypes=12, original_max_haplotypes=12)
    out = engine.select_haplotypes(batch, vset, sel, opt)
    for w in range(batch.n_windows):
        want = S.select_haplotypes(S.window_from_batch(batch, vset, w), 12, 12, 8, 1, 30, opt=opt)
        n = int(out["n_sel"][w])
        assert [int(m) for m in out["sel_mask"][w, :n]] == cases.masks_of([s for s, _ in want]), w
        np.testing.assert_allclose(out["sel_score"][w, :n], [s for _, s in want], rtol=RTOL_TIGHT)


def test_n1_select_schedules_agree(engine, monkeypatch):
    """The two-group pipeline (batches >= 2048 windows; forced here) and the plain round-by-round schedule (no fused
    prologue) return bit for bit what the default schedule returns: same masks, same order, same scores."""
    b1, v1 = synth.make_select_batch(160, n_vars=8, n_reads=32)
    b2, v2 = synth.make_select_batch(90, n_vars=6, n_reads=32, window_offset=160)
    sel = _abi.PlbSelectOptions.default(max_haplotypes=20, original_max_haplotypes=24)
    for b, v in ((b1, v1), (b2, v2)):
        base = engine.select_haplotypes(b, v, sel)
        for env in ({"PLB_SELECT_GROUPS": "2"}, {"PLB_SELECT_NO_PROLOGUE": "1"}, {"PLB_SELECT_GROUPS": "2", "PLB_SELECT_NO_PROLOGUE": "1"}):
            for k in ("PLB_SELECT_GROUPS", "PLB_SELECT_NO_PROLOGUE"):
                monkeypatch.delenv(k, raising=False)
            for k, val in env.items():
                monkeypatch.setenv(k, val)
            got = engine.select_haplotypes(b, v, sel)
            for key in ("n_sel", "sel_mask", "n_scored"):
                assert np.array_equal(got[key], base[key]), (env, key)
            assert np.array_equal(got["sel_score"], base["sel_score"], equal_nan=True), env
        for k in ("PLB_SELECT_GROUPS", "PLB_SELECT_NO_PROLOGUE"):
            monkeypatch.delenv(k, raising=False)
        assert np.all(base["n_sel"] == 19)


def test_n1_select_synth_batch_vs_oracle(engine, oracle):
    """The bench workload's shape (8 variants, 64 reads of 150 bp, 250 bp reference segment, default options: 163 trial
    haplotypes per window in 8 rounds) on 600 windows with mixed variant counts and 2 individuals; 24 windows against
    the oracle, all of them through size-independent properties."""
    from oracle import select_oracle as S
    b1, v1 = synth.make_select_batch(300, n_vars=8, n_individuals=2, n_reads=32)
    b2, v2 = synth.make_select_batch(200, n_vars=6, n_individuals=2, n_reads=32, window_offset=300)
    b3, v3 = synth.make_select_batch(100, n_vars=3, n_individuals=2, n_reads=32, window_offset=500)   # enumerate-all branch
    from platypus_b200.batch import VariantSet, concat_batches
    batch = concat_batches([b2, b1, b3])    # not sorted by variant count: the engine orders the rounds itself
    parts = [v2, v1, v3]
    nv = np.cumsum([0] + [len(v.var_pos) for v in parts])
    na = np.cumsum([0] + [int(v.var_added_off[-1]) for v in parts])
    vset = VariantSet(
        np.concatenate([parts[0].win_var_off] + [v.win_var_off[1:] + nv[i] for i, v in enumerate(parts) if i]).astype(np.int32),
        np.concatenate([v.var_pos for v in parts]), np.concatenate([v.var_n_removed for v in parts]),
        np.concatenate([v.var_n_support for v in parts]),
        np.concatenate([parts[0].var_added_off] + [v.var_added_off[1:] + na[i] for i, v in enumerate(parts) if i]).astype(np.int64),
        np.concatenate([v.var_added[:int(v.var_added_off[-1])] for v in parts] + [np.zeros(1, np.uint8)]))
    out = engine.select_haplotypes(batch, vset)
    st = engine.select_stats()
    assert st["rounds"] == 4 and st["n_filter_windows"] == 500      # one group of windows (< 2048); rounds 0-4 in one launch
    assert np.all(out["n_sel"][:200] == 49) and np.all(out["n_sel"][200:500] == 49) and np.all(out["n_sel"][500:] == 7)
    assert np.all(out["n_scored"][:200] == 1 + 2 + 4 + 8 + 16 + 32) and np.all(out["n_scored"][200:500] == 163)
    for w in range(500):
        sc = out["sel_score"][w, :49]
        assert np.all(np.diff(sc) <= 0)                       # best first
        assert len(set(int(m) for m in out["sel_mask"][w, :49])) == 49
    for w in list(range(0, 200, 25)) + list(range(200, 500, 25)) + [500, 550, 599, 1]:
        want = S.select_haplotypes(S.window_from_batch(batch, vset, w))
        n = int(out["n_sel"][w])
        assert [int(m) for m in out["sel_mask"][w, :n]] == cases.masks_of([s for s, _ in want]), w
        if w < 500:
            np.testing.assert_allclose(out["sel_score"][w, :n], [s for _, s in want], rtol=RTOL_TIGHT)


def test_call_windows_select_build_population_vs_oracle(engine, oracle):
    """The chained flow of callVariantsInWindow on the GPU - selection loop, haplotype construction, window model - against
    the same chain through the oracle: same haplotype lists, same sequences, genotype likelihoods / frequencies / calls /
    variant posteriors within the tolerance."""
    from oracle import select_oracle as S
    from platypus_b200.batch import with_haplotypes
    ref_batch, vset = synth.make_select_batch(24, n_vars=7, n_reads=20, read_len=100, n_individuals=2, seed=4242)
    sel = _abi.PlbSelectOptions.default(max_haplotypes=10, original_max_haplotypes=10)
    sel_out, batch, pop = engine.call_windows(ref_batch, vset, sel)
    hap_off, seqs, masks = [0], [], []
    for w in range(ref_batch.n_windows):
        sw = S.window_from_batch(ref_batch, vset, w)
        chosen = [()] + [s_ for s_, _ in S.select_haplotypes(sw, 10, 10, 8, 1, 30)]
        for s_ in chosen:
            seqs.append(S.build_haplotype(sw.ref_seq, sw.win_start, sw.win_end, sw.hap_start, tuple(sw.vars[i] for i in s_)))
            masks.append(sum(1 << i for i in s_))
        hap_off.append(len(seqs))
    want_batch = with_haplotypes(ref_batch, hap_off, seqs, masks, vset)
    assert np.array_equal(batch.win_hap_off, want_batch.win_hap_off)
    assert np.array_equal(batch.hap_var_mask, want_batch.hap_var_mask)
    assert np.array_equal(batch.hap_seq, want_batch.hap_seq) and np.array_equal(batch.hap_seq_off, want_batch.hap_seq_off)
    want, _, _, _ = oracle.population_run(want_batch, max_haps=pop["max_haps"])
    for k in ("gl", "freq", "em_post", "gof"):
        np.testing.assert_allclose(pop[k], want[k], rtol=RTOL_TIGHT, atol=1e-300, err_msg=k)
    assert np.array_equal(pop["call"], want["call"])
    assert np.array_equal(pop["var_phred"], want["var_phred"])
    assert np.all(sel_out["n_sel"] == 9) and batch.max_haps() == 10


# ---- round 2: packed input, pipelined submit / wait, kernel-side refusals ------------------------------------------

def _pinned_copy(b):
    """The batch with every array in pinned host memory (what plb_population_submit wants)."""
    import dataclasses
    import torch
    kw, keep = {}, []
    for f in b.__dataclass_fields__:
        v = getattr(b, f)
        if isinstance(v, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(v)).pin_memory()
            keep.append(t)
            kw[f] = t.numpy()
    nb = dataclasses.replace(b, **kw)
    nb._keep = []
    nb._pins = keep
    return nb


@pytest.mark.parametrize("mode", [dict(), dict(use_mapq_cap=1), dict(calc_flank_score=1)])
def test_packed_input_is_bit_identical(engine, mode):
    """PLB_SEQ_2BIT batches (2-bit bases + exception lists, the staging format of row N3) give the very same bytes out as
    the ASCII call: edge batches (N, IUPAC, lower case -> exceptions), ragged synth windows, every run-time mode."""
    opt = _abi.PlbOptions.default(**mode)
    binned = synth.make_batch(1200)
    binned.read_qual = np.array([2, 12, 23, 37], np.uint8)[binned.read_qual % 4]       # four quality bins: 4-bit codes
    for b, qbits in ((cases.edge_batch(seed=5), None), (cases.edge_batch(seed=2), None), (binned, 4),
                     (synth.make_batch(1500, read_len_range=(100, 250), hap_len_range=(200, 500)), 6), (synth.make_batch(3000), 6)):
        a = engine.population_run(b, opt=opt, want_ll=True)
        for p in (b.pack(), b.pack(quals=False), b.pack_quals()):
            assert p.input_nbytes() < b.input_nbytes() or p.qual_bits == 0
            assert qbits is None or p.qual_bits in (0, qbits)
            c = engine.population_run(p, opt=opt, want_ll=True)
            for k in ("score", "ll", "gl", "gl_log_max", "gof", "hap_like", "freq", "em_post", "call", "var_phred", "em_iters"):
                assert np.array_equal(a[k], c[k]), k
        assert qbits is None or b.pack().qual_bits == qbits
    # the device-resident path takes packed batches too
    b = cases.edge_batch(seed=3)
    want = engine.population_run(b, opt=opt)
    import torch
    h = engine.upload(b.pack())
    out = {k: torch.zeros(v.shape, dtype=torch.float64 if v.dtype == np.float64 else torch.int32, device="cuda")
           for k, v in want.items() if isinstance(v, np.ndarray)}
    ptrs = {k: v.data_ptr() for k, v in out.items()}
    ptrs["max_haps"] = want["max_haps"]
    engine.run_device(h, ptrs, opt=opt)
    torch.cuda.synchronize()
    engine.last_stats()
    engine.free(h)
    for k in ("gl", "freq", "call", "var_phred"):
        assert np.array_equal(out[k].cpu().numpy(), want[k]), k


def test_submit_wait_two_jobs_in_flight(engine):
    """plb_population_submit / plb_population_wait: batch k+1 is queued while batch k computes (the region loop of
    variantcaller.pyx:566-615).  Results equal the one-call path bit for bit, in any interleaving; a third job in flight
    is refused; synchronous entry points refuse to run under in-flight jobs."""
    from platypus_b200.engine import PlbError
    # 2600 windows: two chunks per queued job, 3300 / 4200: three (tile lists in HBM, one chunk per compute stream)
    batches = [synth.make_batch(n, window_offset=7000 * i) for i, n in enumerate((2600, 3300, 2600, 4200))] + [cases.edge_batch(seed=4)]
    want = [engine.population_run(b) for b in batches]
    pinned = [_pinned_copy(b.pack() if i % 2 else b) for i, b in enumerate(batches)]
    keys = ("gl", "gl_log_max", "gof", "hap_like", "freq", "em_post", "call", "var_phred", "em_iters")
    for rounds in range(2):
        jobs = [engine.population_submit(pinned[0])]
        got = []
        for i in range(1, len(pinned)):
            jobs.append(engine.population_submit(pinned[i]))
            if i == 1:
                with pytest.raises(PlbError) as e:     # PLB_MAX_JOBS = 2
                    engine.population_submit(pinned[2])
                assert e.value.code == _abi.PLB_ERR_ARG
                with pytest.raises(PlbError):
                    engine.upload(batches[0])
            got.append(engine.population_wait(jobs.pop(0)))
        got.append(engine.population_wait(jobs.pop(0)))
        for g, w in zip(got, want):
            for k in keys:
                assert np.array_equal(g[k], w[k]), k
    st = engine.last_stats()
    assert st["n_pairs"] == int(batches[-1].ll_offsets()[-1])
    # an empty batch and waiting twice are harmless
    j = engine.population_submit(synth.make_batch(0), out=engine.alloc_population_out(synth.make_batch(1)))
    engine.population_wait(j)
    engine.population_wait(j)
    engine.population_run(batches[-1])


def test_kernels_refuse_bad_qualities_and_overlong_reads(engine):
    """What the O(slots) host check cannot see is flagged by the kernels and reported by the call: a base quality above
    93 (the reference asserts it, htslibWrapper.pyx:518-519), and - on the device-resident path, which keeps no host
    batch to re-check against the mode - a scored read that does not fit its haplotype (calign.pyx:256-259)."""
    from platypus_b200.engine import PlbError
    b = synth.make_batch(40)
    b.read_qual = b.read_qual.copy()
    b.read_qual[12345] = 94
    with pytest.raises(PlbError) as e:
        engine.population_run(b)
    assert e.value.code == _abi.PLB_ERR_ARG and "quality" in str(e.value)
    b.read_qual[12345] = 93
    engine.population_run(b)
    hap = (b"ACGTTGCA" * 8)[:60]
    r = Read(b"ACGTTGCAAC" * 5, bytes([30] * 50), 100, 150)
    bad = WindowBatch.from_windows([Window(100, 130, 90, [hap], [([], [], [r])])], 1)
    with pytest.raises(PlbError) as e:
        engine.population_run(bad)                     # host path: refused up front
    assert e.value.code == _abi.PLB_ERR_SHAPE
    engine.population_run(synth.make_batch(8))         # the context is still usable


def test_reads_beyond_int16_score_range_take_the_exact_path(engine, oracle):
    """A read whose qualities add up beyond the reference's own score range (15,871 phred, align.c:97) cannot use the
    packed int16 recurrence; it is scored with the 32-bit one and still agrees with the oracle wherever the reference's
    result is defined (scores below the range: a read that matches its haplotype up to a few errors)."""
    rng = random.Random(9)
    wins = []
    for i in range(12):
        L = rng.choice([180, 420, 700])
        hap = cases.random_hap(rng, L + 120)
        at = rng.randint(20, 80)
        read = bytearray(hap[at:at + L])
        for _ in range(rng.randint(0, 4)):
            read[rng.randrange(L)] = ord(rng.choice("ACGT"))
        q = bytes([rng.choice([90, 93, 60])] * L)       # 180 x 93 = 16,740 > 15,871
        r = Read(bytes(read), q, 1000 + at, 1000 + at + L)
        wins.append(Window(1000 + 40, 1000 + 100, 1000, [hap, cases.random_hap(rng, L + 120)[:L + 100] + hap[-20:]], [([r], [], [])]))
    b = WindowBatch.from_windows(wins, 1)
    ll, sc = engine.window_loglik(b)
    ll0, sc0, _ = oracle.window_loglik(b)
    ok = sc0 < 15871                                     # beyond that the reference itself is undefined
    assert ok.sum() >= 12
    assert np.array_equal(sc[ok], sc0[ok])


# ---- the Platypus-shaped objects above the seam (platypus_b200/compat.py) ---------------------------------------------

def _shim_window(case_batch, ref, engine, opts):
    """compat.Haplotype / Variant / WindowReads objects for one l3 population fixture window."""
    from platypus_b200 import compat
    b = case_batch
    vs = [compat.Variant("chr", v[0], v[1], v[2], prior=(0.5 if v[4] is None else v[4])) for v in ref["variants"]]
    H = len(ref["hap_seq"])
    haps = []
    for h in range(H):
        hv = tuple(vs[i] for i, v in enumerate(ref["variants"]) if h in v[6])
        hp = compat.Haplotype("chr", int(b.win_start[0]), int(b.win_end[0]), hv, None, 150, opts, engine,
                              haplotypeSequence=ref["hap_seq"][h])
        hp.hapStart = ref["hap_start"]
        haps.append(hp)
    bufs = []
    for i in range(b.n_individuals):
        s0, s1 = int(b.wi_slot_off[i]), int(b.wi_slot_off[i + 1])
        ng, nb = int(b.wi_n_good[i]), int(b.wi_n_bad[i])
        rd = []
        for s_ in range(s0, s1):
            r = int(b.slot_read[s_])
            o0, o1 = int(b.read_seq_off[r]), int(b.read_seq_off[r + 1])
            rd.append((b.read_seq[o0:o1].tobytes(), b.read_qual[o0:o1].tobytes(), int(b.read_pos[r]), int(b.read_end[r]),
                       int(b.read_mapq[r]), 512 if b.read_qcfail[r] else 0))
        bufs.append(compat.WindowReads(rd[:ng], rd[ng:ng + nb], rd[ng + nb:], sample="s%d" % i))
    return vs, haps, bufs


def test_compat_population_shim_matches_reference(engine, golden_dir):
    """The reference's window loop (variantcaller.pyx:74-141, 566-615) written against platypus_b200.compat:
    Population.reset / setup / call per window, one flush for all of them; every field outputCallToVCF reads equals the
    reference's own Population (tests/golden/l3_pop_ref.npz: 48 windows, 1-8 individuals, all four mode combinations)."""
    import pickle
    from platypus_b200 import compat
    g = np.load(os.path.join(golden_dir, "l3_pop_ref.npz"))
    pops, expect = {}, []
    for seed in range(int(g["n_cases"])):
        c, n_ind, mode, use_em, flat = cases.l3_population_setup(seed)
        key = "p%d_" % seed
        off, hs = g[key + "hap_off"], g[key + "hap"]
        ref = {"hap_seq": [hs[off[k]:off[k + 1]].tobytes() for k in range(len(off) - 1)],
               "hap_start": int(g[key + "hap_start"]), "variants": pickle.loads(g[key + "variants"].tobytes())}
        if flat:
            ref["variants"] = [v[:4] + (None,) + v[5:] for v in ref["variants"]]
        b, phred = cases.l3_population_batch(c, ref, flat)
        pk = (mode, use_em)
        if pk not in pops:
            pops[pk] = compat.Population(compat.Options(HLATyping=mode[0], calculateFlankScore=mode[1], useEMLikelihoods=use_em,
                                                        minPosterior=0), engine, batch_windows=64)
        pop = pops[pk]
        vs, haps, bufs = _shim_window(b, ref, engine, pop.options)
        pop.reset()
        pop.setup(vs, haps, compat.generateAllGenotypesFromHaplotypeList(haps), n_ind, 0, bufs)
        pop.call(100, 1)
        expect.append((pk, {k: g[key + k] for k in ("freq", "gl", "em", "gl_log_max", "gof", "call")}, phred, vs, haps))
    results = {pk: iter(pop.flush()) for pk, pop in pops.items()}
    n = 0
    for pk, want, phred, vs, haps in expect:
        r = next(results[pk])
        nI, G = want["gl"].shape
        np.testing.assert_allclose(r.genotypeLikelihoods, want["gl"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(r.goodnessOfFitValues, want["gof"], rtol=1e-12, atol=0)
        np.testing.assert_allclose(r.frequencies, want["freq"], rtol=1e-9, atol=0)
        np.testing.assert_allclose(r.EMLikelihoods, want["em"], rtol=1e-9, atol=1e-300)
        # genotypes are (hap_i, hap_j) pairs; equal haplotype objects would alias, so compare through identity
        calls = [-1 if gt is None else next(k for k, x in enumerate(r.genotypes) if x[0] is gt[0] and x[1] is gt[1])
                 for gt in r.genotypeCalls]
        assert calls == list(want["call"])
        assert [r.variantPosteriors.get(v) for v in vs] == [float(x) for x in phred]
        assert r.haplotypeIndexes.shape == (G, 2) and list(r.nReads) == [len(x.reads) for x in r.readBuffers]
        n += 1
    assert n == 48
    # limits raise like the reference (cpopulation.pyx:209-225)
    small = compat.Population(compat.Options(maxHaplotypes=2), engine, batch_windows=1)
    with pytest.raises(compat.PlatypusError):
        small.setup(vs, haps + haps + haps, compat.generateAllGenotypesFromHaplotypeList(haps + haps + haps), 1, 0, [None])


def test_compat_haplotype_alignreads_sentinel_and_immediate_population(engine, golden_dir):
    """Haplotype.alignReads returns the likelihood array of chaplotype.pyx:306-377 - good | bad | broken, closed by the
    sentinel 999 - equal to the reference's own values (tests/golden/l3_ref.npz); alignSingleRead skips the QC / overlap
    rule; batch_windows=1 gives the reference's immediate behaviour (results on the population object); haplotype
    sequences built through the GPU constructor equal the reference's."""
    from platypus_b200 import compat
    for k, (b, want) in enumerate(cases.l3_golden_cases(golden_dir)):
        if k >= 12:
            break
        c = cases.l3_window_case(k)
        w_ll, _ = want[(0, 0)]
        H, T = w_ll.shape
        opts = compat.Options()
        genome = c["genome"]

        class Ref:
            def getSequence(self, name, a, e):
                return genome[a:e]
        all_vs = {}
        haps = []
        for hv in c["hap_variants"]:
            vt = tuple(all_vs.setdefault(v, compat.Variant("chr", v[0], v[1], v[2])) for v in hv)
            haps.append(compat.Haplotype("chr", c["win_start"], c["win_end"], vt, Ref(), c["max_read_len"], opts, engine))
        compat.Haplotype.build_sequences(haps, engine)
        seqs = [b.hap_seq[b.hap_seq_off[h]:b.hap_seq_off[h + 1]].tobytes() for h in range(H)]
        assert [h.haplotypeSequence for h in haps] == seqs          # reference-built sequences (l3_ref.npz)
        for h, hp in enumerate(haps):
            arr = hp.alignReads(0, c["good"], c["bad"], c["broken"], 0)
            assert arr[-1] == compat.SENTINEL and len(arr) == T + 1
            np.testing.assert_allclose(arr[:-1], w_ll[h], rtol=1e-12, atol=0)
            assert hp.alignReads(0, [], [], [], 0) is arr                 # cached per individual index
        # a QC-fail good read scores 0 in alignReads but is aligned by alignSingleRead
        qc = [r for r in c["good"] if r[5] & 512 and r[4] > 0]
        if qc:
            assert haps[0].alignSingleRead(qc[0]) < 0.0
        pop = compat.Population(opts, engine, batch_windows=1)
        pop.reset()
        vs = sorted(all_vs.values())
        pop.setup(vs, haps, compat.generateAllGenotypesFromHaplotypeList(haps), 1, 0,
                  [compat.WindowReads(c["good"], c["bad"], c["broken"])])
        pop.call(100, 1)
        direct = engine.population_run(b)
        G = H * (H + 1) // 2
        assert np.array_equal(pop.genotypeLikelihoods, direct["gl"][0, :, :G])
        assert np.array_equal(pop.frequencies, direct["freq"][0, :H]) and len(pop.genotypeCalls) == 1


def test_n3_bam_records_to_packed_batch_to_calls(engine, golden_dir):
    """Rows N3 -> S2/S3 end to end without an ASCII read: the raw records of the reference's test BAM (n3_ref.npz) are
    filtered, trimmed and packed by plb_stage_reads_host, windows are cut by plb_window_slices_host, and the resulting
    PLB_SEQ_2BIT batch over the shared read pool gives bit for bit what BASELINE config 1's ASCII batch gives (whose reads
    went through the Python mirror), i.e. the reference's calign.pyx scores of tests/golden/hla_window_ref.npz."""
    from platypus_b200 import reads as R
    from tests.test_oracle import _n3_records
    g = np.load(os.path.join(golden_dir, "n3_ref.npz"))
    rec = _n3_records(g)
    nb = int(g["n_bam"])
    bam = R.BamRecords(rec.ref_names, *[getattr(rec, k)[:nb] for k in ("ref_id", "pos", "mapq", "flag", "mate_ref_id", "mate_pos", "tlen")],
                       rec.cigar_off[:nb + 1], rec.cigar, rec.seq_off[:nb + 1], rec.nib_off[:nb + 1], rec.nib, rec.qual)
    pool = R.stage_records(bam)
    b_ascii, h = cases.hla_fixture_batch(golden_dir)
    wins, k = [], 0
    for ws, we, hs, nh, ng, nbad in h["windows"]:
        haps = [h["hap"][h["hap_off"][k + j]:h["hap_off"][k + j + 1]].tobytes() for j in range(nh)]
        wins.append((int(ws), int(we), int(hs), haps))
        k += nh
    b = pool.window_batch(wins)
    assert b.seq_format == _abi.PLB_SEQ_2BIT
    assert list(b.wi_n_good) == [int(x[4]) for x in h["windows"]] and list(b.wi_n_bad) == [int(x[5]) for x in h["windows"]]
    assert b.n_reads == nb and b.n_slots == b_ascii.n_slots          # one pool for all windows, slots index into it
    for kw in ({}, dict(use_mapq_cap=1), dict(calc_flank_score=1)):
        opt = _abi.PlbOptions.default(**kw)
        got = engine.population_run(b, opt=opt, want_ll=True)
        want = engine.population_run(b_ascii, opt=opt, want_ll=True)
        for key in ("score", "ll", "gl", "freq", "call", "gof"):
            assert np.array_equal(got[key], want[key]), key
    sc = engine.window_loglik(b)[1]
    scored = sc >= 0          # -1 = short-circuited by the QC-fail / overlap rule (the fixture scores every pair)
    assert scored.sum() > 500 and np.array_equal(sc[scored], h["score_default"][scored])      # the reference's own calign.pyx scores


def test_device_generated_windows_roundtrip_and_parity(engine, oracle):
    """plb_synth_fill_device ("synth-v1d", the device-side generator BASELINE config 5 needs) + plb_batch_download: the
    windows depend on (seed, window id) only - not on how they are chunked - are valid inputs, have the recipe's
    statistics, and the resident results equal the CPU oracle run on the downloaded inputs."""
    import ctypes as C
    import dataclasses
    import torch
    import bench
    old = (bench.C5_IND, bench.C5_READS)
    bench.C5_IND, bench.C5_READS = 5, 12
    try:
        tmpl = bench.c5_template(6)
        tmpl2 = bench.c5_template(2)
    finally:
        bench.C5_IND, bench.C5_READS = old
    h = engine.upload(tmpl)
    engine.synth_fill(h, 1000)
    a = dataclasses.replace(engine.download(h, tmpl), _keep=[])
    a = dataclasses.replace(a, **{f: getattr(a, f).copy() for f in a.__dataclass_fields__ if isinstance(getattr(a, f), np.ndarray)})
    # run the resident batch and compare with the oracle on the downloaded bytes
    W, nI, H = 6, 5, 8
    Gm, V = 36, 14
    out = {"gl": torch.zeros((W, nI, Gm), dtype=torch.float64, device="cuda"), "freq": torch.zeros((W, H), dtype=torch.float64, device="cuda"),
           "em_post": torch.zeros((W, nI, Gm), dtype=torch.float64, device="cuda"), "call": torch.zeros((W, nI), dtype=torch.int32, device="cuda"),
           "var_phred": torch.zeros((W, V), dtype=torch.float64, device="cuda")}
    ptrs = {k: v.data_ptr() for k, v in out.items()}
    ptrs["max_haps"] = H
    engine.run_device(h, ptrs)
    torch.cuda.synchronize()
    engine.last_stats()
    assert engine.lib.plb_validate(C.byref(a.as_struct()), C.byref(_abi.PlbOptions.default()), 0) == 0
    want, _, _, _ = oracle.population_run(a, max_haps=H)
    assert np.array_equal(out["call"].cpu().numpy(), want["call"])
    assert np.array_equal(out["var_phred"].cpu().numpy()[:, :want["var_phred"].shape[1]], want["var_phred"])
    np.testing.assert_allclose(out["gl"].cpu().numpy(), want["gl"], rtol=RTOL_TIGHT, atol=1e-300)
    np.testing.assert_allclose(out["freq"].cpu().numpy(), want["freq"], rtol=RTOL_TIGHT)
    engine.free(h)
    # the recipe: haplotypes distinct, reads mostly match their source, variants declared
    hs = a.hap_seq[:6 * 8 * 250].reshape(6, 8, 250)
    for w in range(6):
        assert len({hs[w, g].tobytes() for g in range(8)}) == 8
    assert set(np.unique(a.read_seq[:-1])) <= set(b"ACGT") and a.read_qual[:-1].min() >= 2 and a.read_qual[:-1].max() <= 40
    assert (a.win_n_var >= 7).all() and (a.win_n_var <= 14).all() and (a.hap_var_mask.reshape(6, 8)[:, 0] == 0).all()
    assert 0.80 < (a.read_mapq == 60).mean() < 0.90
    sc = oracle.window_loglik(a)[1].reshape(6, 5, 8, 12)
    assert np.median(sc.min(axis=2)) < 60          # every read has a haplotype it matches up to sequencing errors
    # chunk independence: windows 1002..1003 generated as their own chunk are the same bytes
    h2 = engine.upload(tmpl2)
    engine.synth_fill(h2, 1002)
    b2 = engine.download(h2, tmpl2)
    engine.free(h2)
    n = 5 * 12 * 150
    assert np.array_equal(b2.read_seq[:2 * n], a.read_seq[2 * n:4 * n]) and np.array_equal(b2.read_qual[:2 * n], a.read_qual[2 * n:4 * n])
    assert np.array_equal(b2.hap_seq[:2 * 8 * 250], a.hap_seq[2 * 8 * 250:4 * 8 * 250])
    assert np.array_equal(b2.read_pos, a.read_pos[2 * 60:4 * 60]) and np.array_equal(b2.hap_var_mask, a.hap_var_mask[16:32])


def test_n1_hla_selection_golden_ref(engine, golden_dir):
    """The --HLATyping haplotype selection (getAllHLAHaplotypesInRegion, variantFilter.pyx:655-736) through
    compat.getAllHLAHaplotypesInRegion: haplotype construction (plb_build_haplotypes_host), both scoring passes
    (plb_best_score_haplotypes_host, plb_best_score_genotypes_host) on the GPU, the heap bookkeeping on the host.  The
    returned list equals the one the REFERENCE'S OWN function returned for the same window (tests/golden/n1_hla_ref.npz:
    57-311 known alleles per window, repeated haplotypes, equal-sequence haplotypes), and both score arrays are the
    reference's to 1e-9."""
    from platypus_b200 import compat
    from platypus_b200.batch import Window
    n_filtered = 0
    for g in cases.hla_golden_cases(golden_dir):
        c = cases.hla_window_case(g["seed"])
        fa, variants, ref_hap, bufs, opts = cases.hla_compat_inputs(c, engine)
        haps = compat.getAllHLAHaplotypesInRegion(b"chr", c["win_start"], c["win_end"], fa, opts, variants, ref_hap, bufs)
        assert [variants.index(h.variants[0]) for h in haps] == g["haps"], g["seed"]
        fv = [v for v in variants if v.varSource == compat.FILE_VAR]
        n_filtered += len(fv) > 150
        if g["seed"] % 4 == 1:      # the two score arrays themselves
            hl = [compat.Haplotype(b"chr", c["win_start"], c["win_end"], (v,), fa, c["max_read_len"], opts, engine) for v in fv]
            for k in range(0, len(hl), 64):
                compat.Haplotype.build_sequences(hl[k:k + 64], engine)
            per_ind = [([compat._as_read(r) for r in rb.reads], [], []) for rb in bufs]
            b = WindowBatch.from_windows([Window(c["win_start"], c["win_end"], hl[0].hapStart, [h.haplotypeSequence for h in hl],
                                                 per_ind)], len(per_ind), dedupe_reads=False)
            hs = engine.best_score_haplotypes(b)
            np.testing.assert_allclose(hs, g["hap_score"], rtol=1e-9, atol=0)
            best = max(range(len(hl)), key=lambda k: (hs[k], hl[k].haplotypeSequence))
            gs = engine.best_score_genotypes(b, [best] * len(hl), list(range(len(hl))), c["opts"]["coverage_sampling_level"])
            np.testing.assert_allclose(gs, g["gt_score"], rtol=1e-9, atol=0)
    assert n_filtered >= 15
    # argument checks of the new entry point
    from platypus_b200.engine import PlbError
    b = cases.edge_batch(seed=3)
    with pytest.raises(PlbError) as e:
        engine.best_score_genotypes(b, [0], [b.n_haps - 1], 30)        # haplotypes of two windows
    assert e.value.code == _abi.PLB_ERR_ARG
    with pytest.raises(PlbError):
        engine.best_score_genotypes(b, [0], [1], 0)                    # targetCoverage must be positive
    assert len(engine.best_score_genotypes(b, [], [], 30)) == 0


@pytest.mark.parametrize("n_windows", [2300, 1974])
def test_large_differential_scores_vs_reference_alignc(engine, oracle, n_windows):
    """>= 1 M adversarial (read, haplotype) pairs through the window path, integer scores bit for bit against the CPU
    oracle's mapAndAlignReadToHaplotype (calign.pyx:170-272) with every band alignment done by the REFERENCE'S OWN
    align.c (oracle/_ref/libalign_ref.so, when present) under OpenMP.  The windows (tests/cases.adversarial_batch) put the
    5-op, 6-op (homopolymers >= 40 bp) and 8-op (haplotype N) recurrences, the byte-exact path, the exact tied-vote path,
    reads of 9-600 bp and > 2000 bp and BAM positions off by +-300 on the path the headline benchmark measures."""
    import time
    n_thr = os.cpu_count() or 1
    t0 = time.time()
    b = cases.adversarial_batch(20261017, n_windows)     # 1974: one chunk whose last DP tiles are split (a past failure)
    n_pairs = int(b.ll_offsets()[-1])
    assert n_pairs >= 900000
    t1 = time.time()
    ll, sc = engine.window_loglik(b)
    st = engine.last_stats()
    kind = oracle.use_reference_kernel(True, traceback=False)
    try:
        ll0, sc0, st0 = oracle.window_loglik(b, n_threads=n_thr)
    finally:
        oracle.use_reference_kernel(False)
    t2 = time.time()
    defined = sc0 < 15871                 # beyond the reference's own int16 range its result is undefined (align.c:81, 97)
    assert defined.mean() > 0.999
    bad = np.nonzero((sc != sc0) & defined)[0]
    assert len(bad) == 0, "scores differ at pairs %s: %s vs %s" % (bad[:8], sc[bad[:8]], sc0[bad[:8]])
    np.testing.assert_allclose(ll[defined], ll0[defined], rtol=RTOL_TIGHT, atol=0)
    assert st["n_pairs"] == st0["n_pairs"] and st["n_pairs_scored"] == st0["n_pairs_scored"] > 700000
    assert st["n_anchor_exact"] > 1000            # the exact tied-vote path did run
    flags = {int(x) for x in np.unique(sc0[sc0 >= 0] == 1000000)}
    print("large differential: %d pairs (%d scored, %d band alignments, %d exact-vote pairs), generated in %.1f s, "
          "GPU + oracle (%s align kernel, %d threads) %.1f s" % (n_pairs, st["n_pairs_scored"], st["n_dp"], st["n_anchor_exact"],
                                                                 t1 - t0, "reference" if kind else "oracle", n_thr, t2 - t1), flags)


def test_full_config2_every_window_vs_oracle(engine, oracle):
    """All 10,000 windows of BASELINE config 2 - 5.12 M pairs - against the oracle with the reference's align.c: integer
    scores bit for bit, per-read LL, genotype likelihoods, EM frequencies, calls and posteriors (not a sample)."""
    n_thr = os.cpu_count() or 1
    b = synth.make_batch_parallel(10000)
    got = engine.population_run(b, want_ll=True)
    oracle.use_reference_kernel(True, traceback=False)
    try:
        want, ll0, sc0, st0 = oracle.population_run(b, n_threads=n_thr)
    finally:
        oracle.use_reference_kernel(False)
    assert np.array_equal(got["score"], sc0)
    np.testing.assert_allclose(got["ll"], ll0, rtol=RTOL_TIGHT, atol=0)
    _check_population(got, want)
    assert engine.last_stats()["cells"] == st0["cells"] == synth.algorithmic_cells(b)
