"""CPU tests of the oracle: against the committed golden vectors (produced by the reference's own
align.c / calign.pyx, see tests/golden/make_golden.py), against the compiled reference when
oracle/_ref is present, and against hand-checkable known answers."""
import math
import os
import random

import numpy as np
import pytest

from tests import cases


def _unpack(off, data, i):
    return data[off[i]:off[i + 1]].tobytes()


def test_golden_align_ref(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "align_ref.npz"))
    n = len(g["score"])
    assert n >= 500
    for i in range(n):
        hap = _unpack(g["hap_off"], g["hap"], i)
        go = _unpack(g["hap_off"], g["gap_open"], i)
        read = _unpack(g["read_off"], g["read"], i)
        qual = _unpack(g["read_off"], g["qual"], i)
        assert oracle.band_align(hap, read, qual, go) == int(g["score"][i]), "golden case %d" % i


def test_golden_calign_ref(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "calign_ref.npz"))
    n = len(g["score"])
    for i in range(n):
        hap = _unpack(g["hap_off"], g["hap"], i)
        read = _unpack(g["read_off"], g["read"], i)
        qual = _unpack(g["read_off"], g["qual"], i)
        s, _ = oracle.map_and_align(read, qual, int(g["read_start"][i]), int(g["hap_start"][i]), hap)
        assert s == int(g["score"][i]), "golden case %d" % i


def test_golden_window_restated(oracle, golden_dir):
    g = np.load(os.path.join(golden_dir, "window_restated.npz"))
    batch = cases.edge_batch(seed=5)
    arrs, ll, sc, _ = oracle.population_run(batch)
    assert np.array_equal(sc, g["score"])
    np.testing.assert_allclose(ll, g["ll"], rtol=1e-12, atol=0)
    for k in ("gl", "gof", "freq", "em_post", "var_phred", "hap_like"):
        np.testing.assert_allclose(arrs[k], g[k], rtol=1e-9, atol=1e-300, err_msg=k)
    assert np.array_equal(arrs["call"], g["call"])
    assert np.array_equal(arrs["em_iters"], g["em_iters"])


def test_vs_ref_alignc_fuzz(oracle):
    if oracle.ref_align_lib() is None:
        pytest.skip("oracle/_ref/libalign_ref.so not built (no reference checkout)")
    rng = random.Random(101)
    for i in range(1500):
        hap, go, read, qual = cases.random_alignment_case(rng, i)
        want = oracle.ref_fast_align(hap, read, qual, go, traceback=bool(i & 1))
        assert oracle.band_align(hap, read, qual, go) == want, "case %d" % i


def test_vs_ref_calign_fuzz(oracle):
    cw = oracle.ref_calign()
    if cw is None:
        pytest.skip("oracle/_ref/calign not built (no reference checkout)")
    rng = random.Random(102)
    for i in range(800):
        hap, read, qual, rs, hs = cases.random_mapping_case(rng, i)
        go = oracle.gap_open(hap)
        want = cw.map_and_align(read, qual, rs, hs, hap, go, 3, 2, 1, 0)
        got, _ = oracle.map_and_align(read, qual, rs, hs, hap, go)
        assert got == want, "case %d" % i


def test_reference_kernel_inside_oracle_agrees(oracle):
    """The 'reference' CPU baseline (oracle anchoring + align.c DP) equals the pure restatement."""
    if not oracle.use_reference_kernel(True):
        pytest.skip("oracle/_ref/libalign_ref.so not built")
    try:
        from platypus_b200 import synth
        b = synth.make_batch(6)
        ll_ref, sc_ref, _ = oracle.window_loglik(b)
    finally:
        oracle.use_reference_kernel(False)
    ll, sc, _ = oracle.window_loglik(b)
    assert np.array_equal(sc, sc_ref)
    assert np.array_equal(ll, ll_ref)


def test_known_answers(oracle):
    """SURVEY §8c spot checks on the reference kernel: exact -> 0, one Q30 mismatch -> 30,
    2 bp deletion -> 43 (= 40 + 3), 2 bp insertion -> 47 (= 40 + 2 + 3 + 2)."""
    rng = random.Random(7)
    hap = cases._rand_seq(rng, 80)
    # avoid accidental homopolymers influencing the hand computation: constant gap-open 40
    go = bytes([40] * 81)
    for off in range(16):
        read = hap[off:off + 50]
        assert oracle.band_align(hap, read, bytes([30] * 50), go) == 0
    read = bytearray(hap[8:58])
    read[20] = ord("A") if read[20] != ord("A") else ord("C")
    assert oracle.band_align(hap, bytes(read), bytes([30] * 50), go) == 30
    assert oracle.band_align(hap, bytes(read), bytes([7] * 50), go) == 7
    hap2 = b"ACGTTGCAAGGCTTAGCCATGATCGGATACCGTTAGCAATGCGTACGATTGCAGTCAGGCTAACGTTGACCATGCAAGT"
    base = hap2[8:60]
    dele = base[:25] + base[27:]           # read lacks 2 haplotype bases
    assert oracle.band_align(hap2, dele, bytes([30] * len(dele)), bytes([40] * 90)) == 43
    ins = base[:25] + b"TT" + base[25:48]  # read has 2 extra bases
    assert oracle.band_align(hap2, ins, bytes([30] * len(ins)), bytes([40] * 90)) == 47


def test_homopolymer_table_formula(oracle):
    """chaplotype.pyx:64-67 evaluated; gap_open() of a run reproduces it."""
    errs = [2.9e-5, 2.9e-5, 2.9e-5, 2.9e-5, 4.3e-5, 1.1e-4, 2.4e-4, 5.7e-4, 1.0e-3, 1.4e-3] + \
           [1.4e-3 + 4.3e-4 * (n - 10) for n in range(11, 50)]
    table = [int(33.5 + 10 * math.log((i + 1) * q) / math.log(0.1)) - 33 for i, q in enumerate(errs)]
    assert len(table) == 49 and table[0] == 45 and table[-1] == 1
    go = oracle.gap_open(b"C" + b"A" * 60 + b"G")
    assert go[-1] == 0 and go[-2] == table[0]
    # position of the last A has run 0, the one before run 1, ... saturating at 48
    for k in range(60):
        assert go[60 - k] == table[min(k, 48)]
    assert go[0] == table[0]
    # N never continues a run
    go = oracle.gap_open(b"ANNNA")
    assert list(go) == [table[0]] * 5 + [0]


def test_hash_aliases(oracle):
    """calign.pyx:61-76: A->1 C->3 G->2 T->0, N aliases G, case-insensitive."""
    assert oracle.kmer_hash(b"AAAAAAA") == int("1" * 7, 4)
    assert oracle.kmer_hash(b"ACGTACG") == int("1320132", 4)
    assert oracle.kmer_hash(b"NNNNNNN") == oracle.kmer_hash(b"GGGGGGG")
    assert oracle.kmer_hash(b"acgtacg") == oracle.kmer_hash(b"ACGTACG")


def test_score_to_ll(oracle):
    assert oracle.score_to_ll(0, 0) == -300.0
    assert oracle.score_to_ll(1000000, 60) == -300.0
    v = oracle.score_to_ll(30, 60)
    assert abs(v - (-0.23025850929940459 * 30 + math.log(1 - 1e-6))) < 1e-12


def test_genotype_mix_branches(oracle):
    import ctypes as C
    L = oracle.lib()
    a = np.array([-1.0, -10.0, -2.0, -2.0005, 0.0], np.float64)
    b = np.array([-5.0, -9.0, -2.5, -2.0, 0.0], np.float64)
    gof = C.c_double()
    v = L.plo_genotype_loglik(a.ctypes.data, b.ctypes.data, 5, 4, 0, C.byref(gof), None, None)
    want = (math.log(0.5) + -1.0) + math.log(0.5 * (math.exp(-10) + math.exp(-9))) + \
        math.log(0.5 * (math.exp(-2) + math.exp(-2.5))) + -2.0005 + 0.0
    assert abs(v - want) < 1e-12
    v_h = L.plo_genotype_loglik(a.ctypes.data, a.ctypes.data, 5, 4, 1, C.byref(gof), None, None)
    assert abs(v_h - a.sum()) < 1e-12
    assert abs(gof.value - (-10 * 0.43429448190325182 * a.sum() / 4)) < 1e-9


# ---- scope row a2 / N2: traceback, calculateFlankScore, HLA map-qual cap ------------------------

def test_golden_align_tb_ref(oracle, golden_dir):
    """Traceback rows, firstpos and calculateFlankScore against the reference's align.c outputs."""
    g = np.load(os.path.join(golden_dir, "align_tb_ref.npz"))
    n = len(g["score"])
    assert n >= 300
    for i in range(n):
        hap = _unpack(g["hap_off"], g["hap"], i)
        go = _unpack(g["hap_off"], g["gap_open"], i)
        read = _unpack(g["read_off"], g["read"], i)
        qual = _unpack(g["read_off"], g["qual"], i)
        a1 = _unpack(g["aln_off"], g["aln1"], i)
        a2 = _unpack(g["aln_off"], g["aln2"], i)
        s, m1, m2, fp = oracle.band_align_tb(hap, read, qual, go)
        assert (s, m1, m2, fp) == (int(g["score"][i]), a1, a2, int(g["firstpos"][i])), "golden case %d" % i
        fl = int(g["hap_flank"][i])
        assert oracle.flank_score(len(hap), fl, qual, go, fp, m1, m2) == int(g["flank_score"][i]), "case %d" % i
        # the one-pass formulation the CUDA path uses
        assert oracle.band_align_flank(hap, read, qual, go, 0, fl) == (s, int(g["flank_score"][i])), "case %d" % i


def test_golden_calign_modes_ref(oracle, golden_dir):
    """mapAndAlignReadToHaplotype with doCalculateFlankScore=1 and with clipped reads voting through
    the unclipped read's hashes (HLA mode) against the reference's calign.pyx outputs."""
    g = np.load(os.path.join(golden_dir, "calign_modes_ref.npz"))
    n = len(g["score"])
    for i in range(n):
        hap = _unpack(g["hap_off"], g["hap"], i)
        read = _unpack(g["read_off"], g["read"], i)
        qual = _unpack(g["read_off"], g["qual"], i)
        hread = _unpack(g["hash_read_off"], g["hash_read"], i)
        s, _ = oracle.map_and_align_ex(read, qual, int(g["read_start"][i]), int(g["hap_start"][i]), hap, None,
                                       int(g["hap_flank"][i]), int(g["do_flank"][i]), hread)
        assert s == int(g["score"][i]), "golden case %d" % i


def test_vs_ref_traceback_and_flank_fuzz(oracle):
    if oracle.ref_align_lib() is None:
        pytest.skip("oracle/_ref/libalign_ref.so not built (no reference checkout)")
    rng = random.Random(201)
    for i in range(1500):
        hap, go, read, qual = cases.random_alignment_case(rng, i)
        want = oracle.ref_fast_align_tb(hap, read, qual, go)
        got = oracle.band_align_tb(hap, read, qual, go)
        assert got == want, "case %d" % i
        fl = rng.randint(1, max(1, len(hap) // 2))
        f = oracle.ref_flank_score(len(hap), fl, qual, go, want[3], want[1], want[2])
        assert oracle.flank_score(len(hap), fl, qual, go, got[3], got[1], got[2]) == f
        assert oracle.band_align_flank(hap, read, qual, go, 0, fl) == (want[0], f)


def test_vs_ref_calign_modes_fuzz(oracle):
    cw = oracle.ref_calign()
    if cw is None:
        pytest.skip("oracle/_ref/calign not built (no reference checkout)")
    rng = random.Random(202)
    for i in range(700):
        hap, read, qual, rs, hs = cases.random_mapping_case(rng, i)
        go = oracle.gap_open(hap)
        fl = rng.randint(1, max(1, len(hap) // 2))
        hread = read
        if i % 2 and len(read) > 30:
            o1, o2 = rng.choice([0, 1, 3, 10]), rng.choice([0, 2, 7])
            read, qual, rs = read[o1:len(read) - o2], qual[o1:len(qual) - o2], rs + o1
        do = int(i % 3 != 0)
        want = cw.map_and_align(read, qual, rs, hs, hap, go, 3, 2, fl, do, 0, hread)
        got, _ = oracle.map_and_align_ex(read, qual, rs, hs, hap, go, fl, do, hread)
        assert got == want, "case %d" % i


def test_score_to_ll_hla(oracle):
    m = -0.23025850929940459
    # below the threshold: standard transform, capped by the log-probability of a wrong mapping
    assert abs(oracle.score_to_ll_hla(10, 60) - (m * 10 + math.log(1 - 1e-6))) < 1e-12
    assert oracle.score_to_ll_hla(100, 20) == m * 20
    assert oracle.score_to_ll_hla(0, 0) == 0.0                 # log(1 - 1) = -inf loses against the cap 0
    # above: mLTOT * (99 + 2*sqrt(score - 99)), chaplotype.pyx:668-672
    assert abs(oracle.score_to_ll_hla(199, 60) - max(m * 60, m * (99 + 2 * math.sqrt(100)))) < 1e-12
    assert abs(oracle.score_to_ll_hla(101, 250) - m * (99 + 2 * math.sqrt(2))) < 1e-12
    assert oracle.score_to_ll_hla(101, 93) == m * 93          # the cap wins


def test_golden_window_modes_restated(oracle, golden_dir):
    from platypus_b200 import _abi
    g = np.load(os.path.join(golden_dir, "window_modes_restated.npz"))
    for name, kw in (("flank", dict(calc_flank_score=1)), ("hla", dict(use_mapq_cap=1)),
                     ("both", dict(calc_flank_score=1, use_mapq_cap=1))):
        batch = cases.edge_batch(seed=5, overhang=True)
        arrs, ll, sc, _ = oracle.population_run(batch, _abi.PlbOptions.default(**kw))
        assert np.array_equal(sc, g[name + "_score"]), name
        np.testing.assert_allclose(ll, g[name + "_ll"], rtol=1e-12, atol=0, err_msg=name)
        np.testing.assert_allclose(arrs["gl"], g[name + "_gl"], rtol=1e-9, atol=1e-300, err_msg=name)
        assert np.array_equal(arrs["call"], g[name + "_call"]), name


def test_reference_kernel_inside_oracle_agrees_flank_mode(oracle):
    """Window path in flank mode: reference align.c + calculateFlankScore vs the restatement."""
    if not oracle.use_reference_kernel(True):
        pytest.skip("oracle/_ref/libalign_ref.so not built")
    from platypus_b200 import _abi
    opt = _abi.PlbOptions.default(calc_flank_score=1)
    try:
        b = cases.edge_batch(seed=8, overhang=True)
        ll_ref, sc_ref, _ = oracle.window_loglik(b, opt)
    finally:
        oracle.use_reference_kernel(False)
    ll, sc, _ = oracle.window_loglik(b, opt)
    assert np.array_equal(sc, sc_ref)
    assert np.array_equal(ll, ll_ref)


# ---- scope row N4: per-site genotype calls -------------------------------------------------------

def test_site_genotypes_known_answer(oracle):
    """computeGenotypeCallAndLikelihoods on a case small enough to do by hand (vcfutils.pyx:163-334):
    two haplotypes (reference, one variant), GL = [0.1, 1.0, 0.2] for genotypes (0,0), (0,1), (1,1)."""
    from platypus_b200.batch import Read, Window, WindowBatch, SiteBatch
    rd = Read(b"A" * 30, bytes([30] * 30), 100, 130)
    w = Window(100, 140, 50, [b"ACGT" * 30, b"ACGA" * 30], [([rd], [], [])], hap_var_mask=[0, 1], var_prior=[1e-3])
    b = WindowBatch.from_windows([w], 1)
    pop = oracle.alloc_population_out(b)
    pop["gl"][0, 0, :3] = [0.1, 1.0, 0.2]
    pop["gof"][0, :3, 0] = [7.0, 3.0, 5.0]
    pop["freq"][0, :2] = [0.5, 0.5]
    sites = SiteBatch.from_lists(b, [(0, [0], [1, 0])])
    r = oracle.site_genotypes(b, pop, sites)
    np.testing.assert_allclose(r["lik"][0, 0], [0.1, 2.0, 0.2], rtol=1e-15)
    np.testing.assert_allclose(r["post"][0, 0], [2.0 / 2.3, 2.2 / 2.3, 0.1 / 2.3], rtol=1e-14)
    assert r["phred"][0, 0].tolist() == [9, 14, 0]
    assert r["phased"][0, 0].tolist() == [0, 1]          # the variant sits on the second haplotype: "0/1"
    assert r["gt"][0, 0].tolist() == [0, 1]
    assert r["gof"][0, 0] == 3.0
    np.testing.assert_allclose(r["gl_log10"][0, 0], [math.log10(0.05), 0.0, math.log10(0.1)], rtol=1e-14)
    # a weak call falls back to 0/0, a hopeless one to ./. (vcfutils.pyx:518-526)
    pop["gl"][0, 0, :3] = [1.0, 0.2, 0.0]     # ref posterior 1/1.4: phred 5; non-ref 0.4/1.4: phred 1
    r = oracle.site_genotypes(b, pop, sites)
    assert r["phred"][0, 0].tolist() == [5, 1, 5] and r["gt"][0, 0].tolist() == [0, 0]
    pop["gl"][0, 0, :3] = [1.0, 0.5, 0.0]
    r = oracle.site_genotypes(b, pop, sites)
    assert r["phred"][0, 0, 1] < 5 and r["phred"][0, 0, 2] < 5 and r["gt"][0, 0].tolist() == [-1, -1]


def test_site_genotypes_edge_batch_properties(oracle):
    b = cases.edge_batch(seed=5)
    pop, _, _, _ = oracle.population_run(b)
    sites = cases.sites_for_batch(b)
    assert sites.n_sites > 20
    r = oracle.site_genotypes(b, pop, sites)
    nI = b.n_individuals
    for s in range(sites.n_sites):
        w = int(sites.site_win[s])
        for i in range(nI):
            if b.wi_n_good[w * nI + i] == 0:
                assert r["gt"][s, i].tolist() == [-1, -1] and r["phred"][s, i].tolist() == [0, 0, 0]
                continue
            tot = r["lik"][s, i].sum()
            if tot > 0:
                assert abs(r["post"][s, i, 0] - r["lik"][s, i].max() / tot) < 1e-12
                assert 0 <= r["phred"][s, i, 0] <= 99


# ---- BASELINE config 1: the reference's own test BAM -----------------------------------------------

def _check_hla_fixture(score_of_mode, g):
    """score_of_mode(mode kwargs) -> flat score array in PlbLoglikOut layout; golden scores hold every pair,
    ours -1 where the QC-fail / overlap rule skips the read (chaplotype.pyx:343-361)."""
    n_scored = 0
    for name, kw in (("default", {}), ("hla", dict(use_mapq_cap=1)), ("flank", dict(calc_flank_score=1))):
        sc = score_of_mode(kw)
        want = g["score_" + name]
        assert len(sc) == len(want)
        live = sc != -1
        assert np.array_equal(sc[live], want[live]), name
        n_scored += int(live.sum())
    assert n_scored > 1500


def test_config1_hla_bam_windows_vs_reference_calign(oracle, golden_dir):
    """Real reads of test/S55_test_realigned.bam against HLA-A allele haplotypes: the oracle's window path
    (all three modes) equals mapAndAlignReadToHaplotype of the reference's calign.pyx pair by pair."""
    from platypus_b200 import _abi
    b, g = cases.hla_fixture_batch(golden_dir)
    assert b.n_windows == 4 and b.n_reads > 150
    _check_hla_fixture(lambda kw: oracle.window_loglik(b, _abi.PlbOptions.default(**kw))[1], g)


def test_read_staging_rules():
    """checkAndTrimRead / setWindowPointers mirrors (cwindow.pyx:332-481, 208-236) on hand-made reads."""
    from platypus_b200 import reads as R
    def mk(pos, flag=0, mapq=60, qual=None, cigar=None, isize=0, mpos=0, n=40):
        q = bytearray(qual if qual is not None else [30] * n)
        return R.AlignedRead(b"A" * n, q, cigar or [(0, n)], 0, pos, pos + n, mapq, flag, 0, mpos, isize)
    opt, cnt = R.ReadFilterOptions(), {}
    assert not R.check_and_trim_read(mk(10, mapq=5), None, opt, cnt) and cnt == {"low_map_qual": 1}
    r = mk(10, qual=[30] * 15 + [3] * 25)
    assert not R.check_and_trim_read(r, None, opt, cnt) and (r.flag & R.F_QCFAIL)          # < 20 good bases
    r = mk(10, flag=R.F_PAIRED | R.F_MATE_UNMAPPED)
    assert not R.check_and_trim_read(r, None, opt, cnt) and not (r.flag & R.F_QCFAIL)      # broken pair: no QC flag
    a, b2 = mk(10), mk(10)
    assert R.check_and_trim_read(a, None, opt, cnt) and not R.check_and_trim_read(b2, a, opt, cnt)   # duplicate
    r = mk(10, qual=[30] * 36 + [4, 30, 2, 1])       # forward read: tail trimmed until a base with q >= 5
    assert R.check_and_trim_read(r, None, opt, cnt) and list(r.qual[36:]) == [4, 30, 0, 0]
    r = mk(10, flag=R.F_REVERSE, qual=[1, 2, 30] + [30] * 37)
    assert R.check_and_trim_read(r, None, opt, cnt) and list(r.qual[:3]) == [0, 0, 30]
    r = mk(10, cigar=[(4, 5), (0, 30), (4, 5)])      # soft clips -> quality 0
    assert R.check_and_trim_read(r, None, opt, cnt) and list(r.qual[:5]) == [0] * 5 and list(r.qual[35:]) == [0] * 5
    r = mk(10, flag=R.F_PAIRED | R.F_PROPER | R.F_MATE_REVERSE, isize=60, mpos=30)   # overlapping mates
    assert R.check_and_trim_read(r, None, opt, cnt) and list(r.qual[-21:]) == [0] * 21 and r.qual[-22] == 30
    rs = [mk(p) for p in (0, 30, 60, 90, 120)]
    got = R.window_slice(rs, 65, 95)                 # first read with pos >= 65 - 40, dropping reads ending <= 65
    assert [x.pos for x in got] == [30, 60, 90]


def test_read_staging_on_reference_bam():
    ref = os.environ.get("PLATYPUS_REFERENCE", "/root/reference")
    bam = os.path.join(ref, "test", "S55_test_realigned.bam")
    if not os.path.exists(bam):
        pytest.skip("reference checkout not present")
    from platypus_b200 import reads as R
    refs, recs = R.decode_bam(bam)
    assert refs[recs[0].chrom_id] == "6" and len(recs) == 2115              # 2115 reads on contig 6
    assert max(r.rlen for r in recs) == 251
    assert all(0 <= q <= 93 for r in recs[:200] for q in r.qual)
    buf = R.ReadBuffer()
    for r in recs:
        buf.add(r)
    assert len(buf.reads) + len(buf.bad_reads) == 2115 and len(buf.reads) > 1500


# ---- L3 pinned: the reference's own Haplotype / DiploidGenotype classes -------------------------------

def test_golden_l3_ref(oracle, golden_dir):
    """Per-read log-likelihoods (Haplotype.alignReads), genotype log-likelihoods, GOF and hapLike
    (DiploidGenotype.calculateDataLikelihood) against outputs of the reference's own chaplotype.pyx /
    cgenotype.pyx, in default, HLA, flank and HLA+flank mode."""
    from platypus_b200 import _abi
    n = 0
    for b, want in cases.l3_golden_cases(golden_dir):
        for (hla, flank), (w_ll, w_geno) in want.items():
            arrs, ll, sc, _ = oracle.population_run(b, _abi.PlbOptions.default(use_mapq_cap=hla, calc_flank_score=flank))
            cases.check_l3(ll, arrs, w_ll, w_geno)
            n += w_ll.size
    assert n > 9000


def test_vs_ref_l3_fuzz(oracle):
    W = oracle.ref_l3()
    if W is None:
        pytest.skip("oracle/_ref/l3_ref_wrap not built (no reference checkout)")
    from platypus_b200 import _abi
    for seed in range(1000, 1120):
        c = cases.l3_window_case(seed)
        hla, flank = cases.L3_MODES[seed % 4]
        r = W.window_likelihoods(c["genome"], c["win_start"], c["win_end"], c["hap_variants"], c["good"], c["bad"],
                                 c["broken"], c["max_read_len"], hla, flank)
        b = cases.l3_case_batch(c, r["hap_seq"], r["hap_start"])
        arrs, ll, sc, _ = oracle.population_run(b, _abi.PlbOptions.default(use_mapq_cap=hla, calc_flank_score=flank))
        cases.check_l3(ll, arrs, np.array(r["ll"]), np.array([g[2:] for g in r["genotypes"]]))


def test_golden_l3_population_ref(oracle, golden_dir):
    """Rescaled genotype likelihoods, EM frequencies, EM genotype posteriors, genotype calls and variant posteriors
    against outputs of the reference's own Population class (cpopulation.pyx setup() + call())."""
    from platypus_b200 import _abi
    n = 0
    for b, want, use_em, (hla, flank) in cases.l3_pop_golden_cases(golden_dir):
        opt = _abi.PlbOptions.default(use_mapq_cap=hla, calc_flank_score=flank, use_em_likelihoods=use_em)
        arrs, _, _, _ = oracle.population_run(b, opt)
        cases.check_l3_pop(arrs, want)
        n += 1
    assert n >= 40


@pytest.mark.filterwarnings("ignore::pytest.PytestUnraisableExceptionWarning")   # the reference's indel PRIOR model
def test_vs_ref_population_fuzz(oracle):                                          # (outside the path) fails under Python 3
    W = oracle.ref_l3()
    if W is None:
        pytest.skip("oracle/_ref/l3_ref_wrap not built (no reference checkout)")
    from platypus_b200 import _abi
    for seed in range(2000, 2060):
        c, n_ind, (hla, flank), use_em, flat = cases.l3_population_setup(seed)
        r = W.population(c["genome"], c["win_start"], c["win_end"], c["hap_variants"], c["per_ind"], c["max_read_len"], hla,
                         flank, use_em)
        b, phred = cases.l3_population_batch(c, r, flat)
        want = {k: np.array(r[k], np.float64) for k in ("freq", "gl", "em", "gl_log_max", "gof")}
        want["call"] = np.array(r["call"], np.int32)
        want["var_phred"] = phred
        opt = _abi.PlbOptions.default(use_mapq_cap=hla, calc_flank_score=flank, use_em_likelihoods=use_em)
        arrs, _, _, _ = oracle.population_run(b, opt)
        cases.check_l3_pop(arrs, want)


def test_golden_n4_ref(oracle, golden_dir):
    """Per-site genotype calls against outputs of the reference's own computeGenotypeCallAndLikelihoods."""
    g = np.load(os.path.join(golden_dir, "n4_ref.npz"))
    for k, (b, sites) in enumerate(cases.n4_cases()):
        pop, _, _, _ = oracle.population_run(b)
        cases.check_n4(oracle.site_genotypes(b, pop, sites), g, k)


def test_vs_ref_n4_fuzz(oracle):
    if oracle.ref_n4() is None:
        pytest.skip("oracle/_ref/n4_ref not built (no reference checkout)")
    for n_ind, seed in ((2, 21), (26, 22)):
        b = cases.edge_batch(seed=seed, n_windows=10, n_individuals=n_ind)
        pop, _, _, _ = oracle.population_run(b)
        sites = cases.sites_for_batch(b, seed=seed)
        ours = oracle.site_genotypes(b, pop, sites)
        for (s, i), (p1, p2, liks, gp, npo, rp, gf) in oracle.ref_site_genotypes(b, pop, sites).items():
            assert ours["phased"][s, i].tolist() == [p1, p2]
            assert np.array_equal(ours["lik"][s, i, :len(liks)], np.array(liks))
            assert np.array_equal(ours["post"][s, i], np.array([gp, npo, rp]), equal_nan=True) and ours["gof"][s, i] == gf


def test_vs_ref_classes_on_synth_workload(oracle):
    """The reference's own classes on config-2 shaped synthetic windows (the path bench.py --impl reference times)
    give exactly the oracle's genotype likelihoods, frequencies and calls."""
    W = oracle.ref_l3()
    if W is None:
        pytest.skip("oracle/_ref/l3_ref_wrap not built (no reference checkout)")
    import bench
    from platypus_b200 import synth
    b = synth.make_batch(12)
    arrs, _, _, _ = oracle.population_run(b)
    for w, a in enumerate(bench._window_args(b, b.n_windows)):
        r = W.population_seq(*a)
        assert np.array_equal(arrs["gl"][w, 0, :36], np.array(r["gl"][0]))
        assert np.array_equal(arrs["freq"][w, :8], np.array(r["freq"]))
        assert arrs["gl_log_max"][w, 0] == r["gl_log_max"][0] and list(arrs["call"][w]) == r["call"]


# ---- N1: haplotype construction + selection loop ------------------------------------------------------------------

def _n1_oracle_run(c, ref_seq, hap_start):
    from oracle import select_oracle as S
    o = c["opts"]
    w = cases.n1_select_window(c, ref_seq, hap_start)
    tr = []
    got = S.select_haplotypes(w, o["max_haplotypes"], o["original_max_haplotypes"], o["max_variants"], o["filter_by_coverage"],
                              o["coverage_sampling_level"], trace=tr)
    sel = [g[0] for g in got]
    seqs = [S.build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start, tuple(w.vars[i] for i in s)) for s in sel]
    return sel, seqs, [vs for rnd in tr for (vs, _) in rnd], [s for rnd in tr for (_, s) in rnd]


def test_golden_n1_ref(oracle, golden_dir):
    """Selection oracle vs the reference's getFilteredHaplotypes / computeBestScoreForGenotype / Haplotype outputs."""
    from oracle.select_oracle import best_score_haplotypes as S_best
    gold = cases.n1_golden_cases(golden_dir)
    assert len(gold) >= 60
    n_filter = n_tied = 0
    for g in gold:
        c = cases.n1_window_case(g["seed"], g["drop"])
        sel, seqs, sets, scores = _n1_oracle_run(c, g["ref_seq"], g["hap_start"])
        assert cases.masks_of(sel) == g["sel_mask"], g["seed"]
        assert seqs == g["hap_seqs"], g["seed"]
        assert cases.masks_of(sets) == g["trial_mask"], g["seed"]
        assert scores == list(g["trial_score"]), g["seed"]          # bit for bit: same libm, same order of additions
        w = cases.n1_select_window(c, g["ref_seq"], g["hap_start"])
        hs = S_best(w, [()] + [tuple(w.vars[i] for i in s_) for s_ in sel])
        assert hs == [g["ref_hap_score"]] + list(g["hap_score"]), g["seed"]      # computeBestScoreForHaplotype, bit for bit
        n_filter += bool(sets)
        n_tied += len(scores) - len(set(scores))
    assert n_filter >= 40 and n_tied >= 100    # the fixture exercises the heap rounds and exactly tied scores


def test_vs_ref_n1_fuzz(oracle):
    ref = oracle.ref_l3()
    if ref is None:
        pytest.skip("reference not available (oracle/_ref)")
    for seed in range(200, 260):
        c = cases.n1_window_case(seed)
        o = c["opts"]
        args = (c["genome"], c["win_start"], c["win_end"], c["variants"], c["per_ind"], c["max_read_len"], o["max_haplotypes"],
                o["original_max_haplotypes"], o["max_variants"], o["filter_by_coverage"], o["coverage_sampling_level"])
        r = ref.select_haplotypes(*args)
        sel, seqs, sets, scores = _n1_oracle_run(c, r["ref_seq"], r["hap_start"])
        assert sel == r["selected"], seed
        assert seqs == r["hap_seqs"], seed
        if sets:
            assert ref.select_haplotypes(*args, 0, sets)["scores"] == scores, seed


# ---- N3 pinned by the reference's own bamReadBuffer (tests/golden/n3_ref.npz) ---------------------------------------------

_N3_OPTS = [
    {},
    {"trim_read_flank": 4, "min_map_qual": 30, "min_base_qual": 25, "min_good_qual_bases": 60},
    {"filter_duplicates": 0, "filter_mate_unmapped": 0, "filter_mate_distant": 0, "filter_small_insert": 0},
    {"trim_overlapping": 0, "trim_adapter": 0, "trim_soft_clipped": 0},
]


def _n3_records(g):
    from platypus_b200 import reads as R
    return R.BamRecords(["6"], *[g["rec_" + k] for k in ("ref_id", "pos", "mapq", "flag", "mate_ref_id", "mate_pos", "tlen", "cigar_off",
                                                          "cigar", "seq_off", "nib_off", "nib", "qual")])


def test_n3_native_staging_golden_ref(golden_dir):
    """plb_stage_reads_host = the reference's ReadIterator.get + addReadToBuffer + checkAndTrimRead (cwindow.pyx:332-481,
    560-595) on all 2115 records of the reference's test BAM and 600 synthetic records that reach the remaining branches,
    for four option sets: list membership, QC-fail flags, trimmed qualities, filter counts - bit for bit; and
    plb_window_slices_host = ReadArray.setWindowPointers (cwindow.pyx:208-236) on 400 windows."""
    import ctypes as C
    from platypus_b200 import _abi, reads as R
    from __graft_entry__ import build
    build()
    from platypus_b200.engine import load_library
    lib = load_library()
    g = np.load(os.path.join(golden_dir, "n3_ref.npz"))
    rec = _n3_records(g)
    assert int(g["n_bam"]) == 2115 and rec.n == 2715
    for s, kw in enumerate(_N3_OPTS):
        pool = R.stage_records(rec, R.ReadFilterOptions(**kw), lib)
        assert np.array_equal(pool.kept, g["kept"])
        assert np.array_equal(pool.good, g["o%d_good" % s]) and np.array_equal(pool.flag, g["o%d_flag" % s]), s
        want_q = g["rec_qual"].copy()
        want_q[g["o%d_zeroed" % s]] = 0
        nb = int(rec.seq_off[-1])
        live = np.repeat(g["kept"], np.diff(rec.seq_off)).astype(bool)
        assert np.array_equal(pool.qual[:nb][live], want_q[:nb][live]), s
        assert pool.counts == list(g["o%d_counts" % s]), (pool.counts, list(g["o%d_counts" % s]))
    pool = R.stage_records(rec, None, lib)
    # bases: the 2-bit pool + exceptions decode to the letters htslib's table gives the nibbles
    nb = int(rec.seq_off[-1])
    i = np.arange(nb)
    back = np.frombuffer(b"ACGT", np.uint8)[(pool.seq2[i >> 2] >> (2 * (i & 3))) & 3].copy()
    back[pool.exc_pos] = pool.exc_chr
    nib = np.zeros(nb, np.uint8)
    for r in range(rec.n):
        b0, L = int(rec.seq_off[r]), int(rec.seq_off[r + 1] - rec.seq_off[r])
        k = np.arange(L)
        nib[b0:b0 + L] = (rec.nib[int(rec.nib_off[r]) + (k >> 1)] >> (4 * (1 - (k & 1)))) & 15
    live = np.repeat(g["kept"], np.diff(rec.seq_off)).astype(bool)
    assert np.array_equal(back[live], np.frombuffer(b"=ACMGRSVTWYHKDBN", np.uint8)[nib][live])
    assert len(pool.exc_pos) == int(np.count_nonzero(~np.isin(nib[live], [1, 2, 4, 8])))
    # window slices of the good and the bad list
    wins, want = g["windows"], g["slices"]
    for col, idx in ((0, pool.good_index()), (2, pool.bad_index())):
        pos, end = np.ascontiguousarray(pool.read_pos[idx]), np.ascontiguousarray(pool.read_end[idx])
        lo, hi = np.zeros(len(wins), np.int32), np.zeros(len(wins), np.int32)
        ws, we = np.ascontiguousarray(wins[:, 0]), np.ascontiguousarray(wins[:, 1])
        assert lib.plb_window_slices_host(len(idx), _abi.ptr(pos), _abi.ptr(end), len(wins), _abi.ptr(ws), _abi.ptr(we),
                                          _abi.ptr(lo), _abi.ptr(hi)) == 0
        assert np.array_equal(lo, want[:, col]) and np.array_equal(hi, want[:, col + 1])
        assert (hi > lo).any()


def test_n3_python_mirror_golden_ref(golden_dir):
    """The Python mirror of the same steps (platypus_b200/reads.py: decode_bam's field derivation, check_and_trim_read,
    window_slice) against the same reference-made fixture."""
    from platypus_b200 import reads as R
    g = np.load(os.path.join(golden_dir, "n3_ref.npz"))
    rec = _n3_records(g)
    nibtab = b"=ACMGRSVTWYHKDBN"
    for s, kw in enumerate(_N3_OPTS[:2]):
        buf = R.ReadBuffer(R.ReadFilterOptions(**kw))
        objs = {}
        for i in range(rec.n):
            b0, b1 = int(rec.seq_off[i]), int(rec.seq_off[i + 1])
            if b1 == b0 or rec.qual[b0] == 0xFF:
                continue
            nb = rec.nib[int(rec.nib_off[i]):int(rec.nib_off[i + 1])]
            seq = bytes(nibtab[(nb[k >> 1] >> (4 * (1 - (k & 1)))) & 15] for k in range(b1 - b0))
            cg = [(int(w) & 15, int(w) >> 4) for w in rec.cigar[int(rec.cigar_off[i]):int(rec.cigar_off[i + 1])]]
            pos = int(rec.pos[i]) - (cg[0][1] if cg and cg[0][0] == 4 else 0)
            ref_len = sum(n for op, n in cg if op in (0, 2, 3, 7, 8))
            r = R.AlignedRead(seq, bytearray(rec.qual[b0:b1].tobytes()), cg, int(rec.ref_id[i]), pos,
                              int(rec.pos[i]) + (ref_len if ref_len > 0 else 1), int(rec.mapq[i]), int(rec.flag[i]),
                              int(rec.mate_ref_id[i]), int(rec.mate_pos[i]), int(rec.tlen[i]))
            buf.add(r)
            objs[i] = r
        good_ids = {id(r) for r in buf.reads}
        want_q = g["rec_qual"].copy()
        want_q[g["o%d_zeroed" % s]] = 0
        for i, r in objs.items():
            assert (id(r) in good_ids) == bool(g["o%d_good" % s][i]), i
            assert r.flag == int(g["o%d_flag" % s][i]), i
            assert bytes(r.qual) == want_q[int(rec.seq_off[i]):int(rec.seq_off[i + 1])].tobytes(), i
        if s == 0:
            for (ws, we), (gl, gh, bl, bh) in zip(g["windows"], g["slices"]):
                assert [id(x) for x in R.window_slice(buf.reads, int(ws), int(we))] == [id(x) for x in buf.reads[gl:gh]]
                assert [id(x) for x in R.window_slice(buf.bad_reads, int(ws), int(we))] == [id(x) for x in buf.bad_reads[bl:bh]]


def test_golden_l3_population_many_samples(oracle, golden_dir):
    """The oracle's Population restatement at BASELINE config 5's sample count: 300 and 2000 individuals against outputs of
    the reference's own cpopulation.pyx (tests/golden/l3_pop_many_ref.npz)."""
    from platypus_b200 import _abi
    n = 0
    for b, want, use_em in cases.l3_pop_many_cases(golden_dir):
        arrs, _, _, _ = oracle.population_run(b, _abi.PlbOptions.default(use_em_likelihoods=use_em))
        cases.check_l3_pop_many(arrs, want)
        n += b.n_individuals
    assert n == 4600


# ---- the --HLATyping haplotype selection (variantFilter.pyx:655-736) ---------------------------------------------------
def _hla_window(case, g):
    c4 = dict(case)
    c4["variants"] = [v[:4] for v in case["variants"]]
    return cases.n1_select_window(c4, g["ref_seq"], g["hap_start"])


def test_hla_selection_oracle_vs_reference_golden(golden_dir):
    """The restated getAllHLAHaplotypesInRegion (oracle/select_oracle.hla_haplotypes: one haplotype per FILE_VAR variant,
    the 150-haplotype short cut, the shared heap of (score, haplotype) tuples over both passes, repeated output) returns
    the reference's own list on every fixture window, and its two scoring functions give the reference's numbers."""
    from oracle import select_oracle as S
    n_filtered = 0
    for g in cases.hla_golden_cases(golden_dir):
        c = cases.hla_window_case(g["seed"])
        w = _hla_window(c, g)
        src = [v[4] for v in c["variants"]]
        got = S.hla_haplotypes(w, src, c["opts"]["original_max_haplotypes"], c["opts"]["coverage_sampling_level"])
        assert got == g["haps"], g["seed"]
        fv = [i for i, s_ in enumerate(src) if s_ == 2]
        n_filtered += len(fv) > 150
        if g["seed"] % 6 == 1:      # the scores themselves (a few windows: every one costs ~300 haplotypes x 100 reads)
            sets = [(w.vars[i],) for i in fv]
            hs = S.best_score_haplotypes(w, sets)
            np.testing.assert_allclose(hs, g["hap_score"], rtol=1e-12, atol=0)
            best = max(zip(hs, [S._HapKey(S.build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start, vs), i)
                                for vs, i in zip(sets, fv)]), key=lambda t: (t[0], t[1].seq))[1]
            pos = fv.index(best.idx)
            gs = S.best_score_genotype_pairs(w, [sets[pos]] * len(sets), sets, c["opts"]["coverage_sampling_level"])
            np.testing.assert_allclose(gs, g["gt_score"], rtol=1e-12, atol=0)
    assert n_filtered >= 15


class _OracleScoringEngine:
    """Stands in for Engine in the CPU test of the HOST logic of compat.getAllHLAHaplotypesInRegion: sequences and scores
    come from the fixture / the oracle, so what is tested is the bookkeeping above the C ABI (no product path uses it)."""

    def __init__(self, window, fixture, file_vars):
        self.w, self.g, self.fv = window, fixture, file_vars

    def build_haplotypes(self, ref_batch, vset, hap_win, hap_mask):
        from oracle import select_oracle as S
        out = []
        for m in hap_mask:
            k = int(m).bit_length() - 1
            p, nrem = int(vset.var_pos[k]), int(vset.var_n_removed[k])
            add = bytes(vset.var_added[int(vset.var_added_off[k]):int(vset.var_added_off[k + 1])])
            v = next(v for v in self.w.vars if v.pos == p and len(v.removed) == nrem and v.added == add)
            out.append(S.build_haplotype(self.w.ref_seq, self.w.win_start, self.w.win_end, self.w.hap_start, (v,)))
        return out

    def best_score_haplotypes(self, batch, opt=None):
        return np.asarray(self.g["hap_score"])

    def best_score_genotypes(self, batch, hap1, hap2, target_coverage=30, opt=None):
        self.best = int(hap1[0])
        return np.asarray(self.g["gt_score"])


def test_compat_hla_selection_host_logic(golden_dir):
    """compat.getAllHLAHaplotypesInRegion - the reference's argument list, Haplotype ordering inside the (score, haplotype)
    tuples, the heap shared by both passes - returns the reference's list when the scores are the reference's."""
    from platypus_b200 import compat
    for g in cases.hla_golden_cases(golden_dir):
        c = cases.hla_window_case(g["seed"])
        w = _hla_window(c, g)
        fv = [i for i, v in enumerate(c["variants"]) if v[4] == 2]
        eng = _OracleScoringEngine(w, g, fv)
        fa, variants, ref_hap, bufs, opts = cases.hla_compat_inputs(c, eng)
        haps = compat.getAllHLAHaplotypesInRegion(b"chr", c["win_start"], c["win_end"], fa, opts, variants, ref_hap, bufs)
        got = [variants.index(h.variants[0]) for h in haps]
        assert got == g["haps"], g["seed"]

