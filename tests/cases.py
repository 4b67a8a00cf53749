"""Deterministic test-case generators shared by the golden-vector script and the tests."""
import os
import random

import numpy as np

from platypus_b200.batch import Read, Window, WindowBatch

ACGT = b"ACGT"


def _rand_seq(rng, n, alphabet=ACGT):
    return bytes(rng.choice(alphabet) for _ in range(n))


def mutate(rng, src: bytes, L: int, sub=0.02, ins=0.008, dele=0.008, n_rate=0.004):
    """Read of length L copied from src with substitutions / indels / N."""
    out = bytearray()
    i = 0
    while len(out) < L:
        u = rng.random()
        if i >= len(src):
            out.append(rng.choice(ACGT))
        elif u < sub:
            out.append(rng.choice(ACGT))
            i += 1
        elif u < sub + ins:
            out.append(rng.choice(ACGT))
        elif u < sub + ins + dele:
            i += rng.randint(1, 3)
        elif u < sub + ins + dele + n_rate:
            out.append(ord("N"))
            i += 1
        else:
            out.append(src[i])
            i += 1
    return bytes(out)


def random_alignment_case(rng, i):
    """(hap_segment >= L+15, gap_open, read, qual) for the L1 kernel."""
    L = rng.choice([1, 2, 7, 8, 9, 10, 15, 16, 17, 20, 24, 31, 37, 64, 100, 101, 150, 151, 250, 251])
    n = L + 15 + rng.randint(0, 30)
    hap = bytearray(_rand_seq(rng, n))
    kind = i % 8
    if kind == 1:  # N's in the haplotype (free match, align.c:175-178)
        for _ in range(rng.randint(1, 6)):
            hap[rng.randrange(n)] = ord("N")
    if kind == 2:  # homopolymer
        p = rng.randrange(max(1, n - 12))
        for k in range(p, min(n, p + rng.randint(4, 12))):
            hap[k] = hap[p]
    if kind == 3:  # IUPAC / lower case bytes: byte-exact compare (align.c:315)
        for _ in range(rng.randint(1, 4)):
            hap[rng.randrange(n)] = rng.choice(b"RYKMacgtn")
    off = rng.randint(0, 15)
    if kind == 4:
        read = _rand_seq(rng, L, b"ACGTN")  # unrelated read
    elif kind == 5:  # leading insertion: read starts with bases absent from the haplotype
        k = rng.randint(1, min(4, L))
        read = (_rand_seq(rng, k) + bytes(hap[off:off + L]))[:L]
    else:
        read = mutate(rng, bytes(hap[off:]), L)
    if kind == 3 and L > 4:  # let a read carry the same odd byte so exact equality matters
        j = rng.randrange(L)
        read = read[:j] + bytes([hap[min(n - 1, off + j)]]) + read[j + 1:]
    qual = bytes(0 if rng.random() < 0.08 else rng.randint(1, 41) for _ in range(L))
    if kind == 6:
        qual = bytes(rng.randint(30, 60) for _ in range(L))  # high qualities, bigger scores
    go = bytes(rng.choice([45, 42, 39, 32, 23, 16, 9, 5, 1]) for _ in range(n + 1))
    return bytes(hap), go, read, qual


def random_hap(rng, n):
    u = rng.random()
    if u < 0.3:  # tandem repeat -> tied votes
        unit = _rand_seq(rng, rng.randint(1, 9))
        h = bytearray((unit * (n // len(unit) + 1))[:n])
        for _ in range(rng.randint(0, 6)):
            h[rng.randrange(n)] = rng.choice(ACGT)
    else:
        h = bytearray(_rand_seq(rng, n))
    if rng.random() < 0.15:
        for _ in range(rng.randint(1, 6)):
            h[rng.randrange(n)] = ord("N")
    if rng.random() < 0.05:
        h[rng.randrange(n)] = rng.choice(b"RYacgt")
    return bytes(h)


def random_mapping_case(rng, i):
    """(hap, read, qual, read_start, hap_start) for mapAndAlignReadToHaplotype."""
    L = rng.choice([3, 7, 8, 9, 12, 30, 50, 75, 100, 150, 200])
    hap_len = L + 16 + rng.randint(0, 400)
    hap = random_hap(rng, hap_len)
    idx = rng.randint(0, hap_len - L)
    if rng.random() < 0.1:
        read = _rand_seq(rng, L, b"ACGTN")
    else:
        read = mutate(rng, hap[idx:], L, sub=0.02, ins=0.005, dele=0.005, n_rate=0.002)
    qual = bytes(rng.randint(0, 40) for _ in range(L))
    hap_start = rng.randint(100, 100000)
    jit = rng.choice([0, 0, 0, rng.randint(-5, 5), rng.randint(-300, 300)])
    if i % 37 == 0:
        jit = -idx - 1  # fallback index exactly -1
    return hap, read, qual, hap_start + idx + jit, hap_start


def edge_batch(seed=5, n_windows=12, n_individuals=3, overhang=False):
    """Small multi-individual batch exercising the edge cases the reference's logic has:
    N / IUPAC bytes, zeroed qualities, QC-fail, overlap < 7, mapq 0, exact matches, tandem repeats,
    reads shorter than 7 / 9, ragged lengths, individuals without reads, bad + broken-mate reads,
    single-haplotype windows, shared reads between windows."""
    rng = random.Random(seed)
    windows = []
    shared = None
    for w in range(n_windows):
        hap_len = rng.choice([120, 200, 260, 333])
        flank = (hap_len - 40) // 2
        hap_start = 5000 + 700 * w
        ws, we = hap_start + flank, hap_start + flank + 40
        ref = bytearray(random_hap(rng, hap_len) if w % 3 == 1 else _rand_seq(rng, hap_len))
        if w % 4 == 2:
            ref[rng.randrange(hap_len)] = ord("N")
        if w == 7:
            ref[flank + 3] = ord("R")  # IUPAC -> general path for the whole window
        H = 1 if w == 5 else rng.choice([2, 3, 4, 6])
        haps = [bytes(ref)]
        masks = [0]
        nvar = 0
        while len(haps) < H:
            h = bytearray(ref)
            m = 0
            for _ in range(rng.randint(1, 2)):
                p = flank + rng.randrange(40)
                kind = rng.random()
                if kind < 0.6:
                    h[p] = rng.choice([c for c in ACGT if c != h[p]])
                elif kind < 0.8:
                    h[p:p] = _rand_seq(rng, rng.randint(1, 3))
                else:
                    del h[p:p + rng.randint(1, 3)]
                m |= 1 << (nvar % 6)
                nvar += 1
            h = bytes(h[:hap_len]) if len(h) >= hap_len else bytes(h) + _rand_seq(rng, hap_len - len(h))
            if h not in haps:
                haps.append(h)
                masks.append(m)
        n_var = min(6, max(1, nvar)) if H > 1 else 0
        per_ind = []
        for i in range(n_individuals):
            if (w + i) % 5 == 4:
                per_ind.append(([], [], []))  # individual without any reads
                continue
            lists = ([], [], [])
            n_reads = rng.randint(1, 14)
            for k in range(n_reads):
                L = rng.choice([5, 7, 8, 9, 20, 36, 50, 75, 76, 100])
                L = min(L, hap_len - 16)
                src = haps[rng.randrange(len(haps))]
                idx = rng.randint(0, hap_len - L - 15)
                u = rng.random()
                if u < 0.12:
                    seq = src[idx:idx + L]  # exact match
                elif u < 0.2:
                    seq = _rand_seq(rng, L, b"ACGTN")
                else:
                    seq = mutate(rng, src[idx:], L)
                qual = bytes(0 if rng.random() < 0.1 else rng.randint(2, 41) for _ in range(L))
                pos = hap_start + idx + rng.choice([0, 0, 0, rng.randint(-5, 5), rng.randint(-60, 60)])
                if overhang and rng.random() < 0.35:
                    # reads hanging over either end of the haplotype: HLA mode clips them
                    # (chaplotype.pyx:647-655; the right-hand clip is measured from startPos + hapLen)
                    if rng.random() < 0.5:
                        k = rng.randint(1, min(25, L - 1))
                        seq = _rand_seq(rng, k) + mutate(rng, src, L - k)
                        pos = hap_start - k
                    else:
                        k = rng.randint(1, min(25, L - 1))
                        pos = hap_start + flank + hap_len - L + k
                        seq = mutate(rng, src[max(0, hap_len - L):], L)
                mapq = rng.choice([60, 60, 60, 37, 20, 3, 0])
                r = Read(seq, qual, pos, pos + L, mapq, qcfail=(rng.random() < 0.08))
                which = 0 if rng.random() < 0.7 else rng.choice([1, 2])
                lists[which].append(r)
            if w % 2 == 1 and shared is not None and i == 0:
                lists[0].append(shared)  # a read object shared with the previous window
            if lists[0]:
                shared = lists[0][0]
            # the no-data rule keys on good reads only (cpopulation.pyx:286-294): make one individual
            # have bad reads but no good ones
            if w == 3 and i == 1:
                lists = ([], lists[0] + lists[1], lists[2])
            per_ind.append(lists)
        windows.append(Window(ws, we, hap_start, haps, per_ind, hap_var_mask=masks,
                              var_prior=[rng.choice([1e-3, 1e-4, 0.5, 3.3e-4]) for _ in range(n_var)]))
    return WindowBatch.from_windows(windows, n_individuals)


def sites_for_batch(batch, seed=3):
    """Reported sites for a WindowBatch (scope row N4): one bi-allelic site per window variant, plus a
    multi-allelic site where a window has two or more variants.  haplotypeIsRefAtThisPos is 0 for the
    haplotypes that carry a variant of the site and, now and then, for one carrying another variant that
    spans the position (vcfutils.pyx:403-417)."""
    from platypus_b200.batch import SiteBatch
    rng = random.Random(seed)
    sites = []
    for w in range(batch.n_windows):
        nv = int(batch.win_n_var[w]) if batch.win_n_var is not None else 0
        h0, h1 = int(batch.win_hap_off[w]), int(batch.win_hap_off[w + 1])
        masks = [int(batch.hap_var_mask[h]) for h in range(h0, h1)]

        def is_ref(vs):
            out = []
            for m in masks:
                r = 0 if any((m >> v) & 1 for v in vs) else 1
                if r and m and rng.random() < 0.15:
                    r = 0   # another variant of this haplotype spans the position
                out.append(r)
            return out
        for v in range(nv):
            sites.append((w, [v], is_ref([v])))
        if nv >= 2:
            vs = rng.sample(range(nv), 2)
            sites.append((w, vs, is_ref(vs)))
        if nv >= 3 and rng.random() < 0.5:
            vs = rng.sample(range(nv), 3)
            sites.append((w, vs, is_ref(vs)))
    return SiteBatch.from_lists(batch, sites, min_posterior=5)


def hla_fixture_batch(golden_dir):
    """BASELINE config 1: the windows of tests/golden/hla_window_ref.npz (real reads of the reference's
    test BAM, HLA-A allele haplotypes; made by tests/golden/make_hla_fixture.py).  Returns (batch, fixture)."""
    import os
    import numpy as np
    g = np.load(os.path.join(golden_dir, "hla_window_ref.npz"))
    windows = []
    h = r = 0
    for ws, we, hs, nh, ng, nb in g["windows"]:
        haps = [g["hap"][g["hap_off"][h + k]:g["hap_off"][h + k + 1]].tobytes() for k in range(nh)]
        rds = []
        for k in range(ng + nb):
            i = r + k
            rds.append(Read(g["read"][g["read_off"][i]:g["read_off"][i + 1]].tobytes(),
                            g["qual"][g["read_off"][i]:g["read_off"][i + 1]].tobytes(), int(g["read_pos"][i]),
                            int(g["read_end"][i]), int(g["read_mapq"][i]), bool(g["read_qcfail"][i])))
        windows.append(Window(int(ws), int(we), int(hs), haps, [(rds[:ng], rds[ng:], [])]))
        h += nh
        r += ng + nb
    return WindowBatch.from_windows(windows, 1), g


L3_MODES = [(0, 0), (1, 0), (0, 1), (1, 1)]   # (use_mapq_cap, calc_flank_score)


def l3_window_case(seed, n_ind=1):
    """One window for the reference's Haplotype / DiploidGenotype classes (oracle/l3_ref_wrap.pyx): a random
    genome, a 40-60 bp window, 2-5 haplotypes made of SNP / insertion / deletion variants (as
    (refPos, removed, added) tuples, sorted, non-overlapping) and good / bad / broken-mate reads drawn from the
    haplotypes with errors, low qualities, mapq 0, QC-fail flags, reads that barely overlap the window and
    reads hanging over the haplotype ends."""
    rng = random.Random(seed)
    genome = _rand_seq(rng, 2400)
    if seed % 3 == 0:   # a homopolymer inside the window (context-dependent gap-open penalties)
        genome = genome[:1210] + bytes([rng.choice(ACGT)]) * rng.randint(6, 14) + genome[1224:]
        genome = genome[:2400]
    ws = 1200
    we = ws + rng.randint(40, 60)
    max_read_len = rng.choice([100, 150])
    n_haps = rng.randint(2, 5)
    hap_variants = [[]]
    while len(hap_variants) < n_haps:
        vs, pos = [], ws + rng.randint(2, 8)
        for _ in range(rng.randint(1, 3)):
            if pos >= we - 4:
                break
            kind = rng.random()
            if kind < 0.6:
                alt = rng.choice([c for c in ACGT if c != genome[pos]])
                vs.append((pos, genome[pos:pos + 1], bytes([alt])))
            elif kind < 0.8:
                vs.append((pos, b"", _rand_seq(rng, rng.randint(1, 3))))
            else:
                k = rng.randint(1, 3)
                vs.append((pos, genome[pos:pos + k], b""))
            pos += rng.randint(5, 15)
        if vs and vs not in hap_variants:
            hap_variants.append(vs)
    flank = min(2 * max_read_len, 500)

    def apply(vs):   # plain substitution, only to draw reads from (the reference builds its own sequences)
        out, cur = bytearray(), ws - flank
        for p, rem, add in vs:
            if rem and add:
                out += genome[cur:p] + add
                cur = p + len(rem)
            elif add:
                out += genome[cur:p + 1] + add
                cur = p + 1
            else:
                out += genome[cur:p]
                cur = p + len(rem)
        out += genome[cur:we + flank]
        return bytes(out)
    seqs = [apply(vs) for vs in hap_variants]

    def reads(n):
        out = []
        for _ in range(n):
            L = rng.choice([30, 60, 75, max_read_len])
            src = rng.choice(seqs)
            u = rng.random()
            if u < 0.75:
                idx = rng.randint(max(0, flank - L + 8), min(len(src) - L - 16, flank + (we - ws) - 8))
            elif u < 0.9:
                idx = rng.randint(0, len(src) - L - 16)            # may not overlap the window at all
            else:
                idx = rng.randint(0, 4)                           # starts at the haplotype's left end
            seq = mutate(rng, src[idx:], L, n_rate=0.003)
            qual = bytes(0 if rng.random() < 0.05 else rng.randint(2, 41) for _ in range(L))
            pos = ws - flank + idx + rng.choice([0, 0, 0, rng.randint(-4, 4), rng.randint(-40, 10)])
            flag = 512 if rng.random() < 0.06 else 0
            out.append((seq, qual, pos, pos + L, rng.choice([60, 60, 60, 40, 23, 5, 0]), flag))
        return out
    first = (reads(rng.randint(4, 14)), reads(rng.randint(0, 4)), reads(rng.randint(0, 3)))
    per_ind = [first]
    for i in range(1, n_ind):   # further individuals (drawn after the first one: the n_ind = 1 stream is unchanged)
        if i % 4 == 2:
            per_ind.append(([], reads(rng.randint(0, 2)), []))          # no good reads: "no data" for the model
        else:
            per_ind.append((reads(rng.randint(1, 10)), reads(rng.randint(0, 2)), reads(rng.randint(0, 1))))
    return dict(genome=genome, win_start=ws, win_end=we, hap_variants=hap_variants, good=first[0], bad=first[1],
                broken=first[2], per_ind=per_ind, max_read_len=max_read_len)


def l3_case_batch(case, hap_seqs, hap_start, masks=None, priors=None):
    """WindowBatch of an l3_window_case (all its individuals), haplotype sequences as the reference built them."""
    def mk(t):
        return Read(t[0], t[1], t[2], t[3], t[4], bool(t[5] & 512))
    per = case["per_ind"]
    w = Window(case["win_start"], case["win_end"], hap_start, list(hap_seqs),
               [([mk(t) for t in g], [mk(t) for t in b], [mk(t) for t in k]) for (g, b, k) in per],
               hap_var_mask=masks, var_prior=priors)
    return WindowBatch.from_windows([w], len(per))


def l3_population_setup(seed):
    """(case, n_ind, (hla, flank), use_em, flat_prior) of the population fixtures."""
    rng = random.Random(seed * 7919 + 13)
    n_ind = rng.choice([1, 2, 3, 5, 8])
    return l3_window_case(seed, n_ind), n_ind, L3_MODES[seed % 4], seed % 2, seed % 3 == 0


def l3_population_batch(case, ref_out, flat_prior):
    """Batch + expected phred posteriors from an oracle/l3_ref_wrap.population() result (or its golden copy):
    variant v of the window = entry v of ref_out["variants"]."""
    H = len(ref_out["hap_seq"])
    masks = [0] * H
    priors, phred = [], []
    for vi, v in enumerate(ref_out["variants"]):
        for h in v[6]:
            masks[h] |= 1 << vi
        use_flat = flat_prior or v[4] is None
        priors.append(0.5 if use_flat else v[4])
        phred.append(v[3] if use_flat else v[5])
    return l3_case_batch(case, ref_out["hap_seq"], ref_out["hap_start"], masks, priors), phred





def l3_golden_cases(golden_dir):
    """Yields (batch, {(hla, flank): (ll [H][T], geno [G][4] = logL, gof, hap1Like, hap2Like)}) for the windows
    of tests/golden/l3_ref.npz (outputs of the reference's own chaplotype.pyx / cgenotype.pyx)."""
    import os
    import numpy as np
    g = np.load(os.path.join(golden_dir, "l3_ref.npz"))
    for seed in range(int(g["n_cases"])):
        c = l3_window_case(seed)
        off, hs = g["c%d_hap_off" % seed], g["c%d_hap" % seed]
        haps = [hs[off[k]:off[k + 1]].tobytes() for k in range(len(off) - 1)]
        b = l3_case_batch(c, haps, int(g["c%d_hap_start" % seed]))
        yield b, {m: (g["c%d_m%d%d_ll" % (seed, m[0], m[1])], g["c%d_m%d%d_geno" % (seed, m[0], m[1])]) for m in L3_MODES}


def check_l3(ll, pop, want_ll, want_geno, rtol=1e-9, exact_ll=False):
    """Per-read log-likelihoods, genotype log-likelihoods (ours are stored rescaled: log(gl) + gl_log_max), GOF and
    hapLike of ONE single-individual window against the reference's values."""
    import numpy as np
    H, T = want_ll.shape
    np.testing.assert_allclose(np.asarray(ll).reshape(H, T), want_ll, rtol=1e-12, atol=0)
    if exact_ll:   # mLTOT * score + log(1 - exp(mLTOT * mapq)): the same two roundings and the same libm value
        assert np.array_equal(np.asarray(ll).reshape(H, T), want_ll)
    G = H * (H + 1) // 2
    logl = np.log(pop["gl"][0, 0, :G]) + pop["gl_log_max"][0, 0]
    np.testing.assert_allclose(logl, want_geno[:, 0], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(pop["gof"][0, :G, 0], want_geno[:, 1], rtol=1e-12, atol=0)
    g = 0
    for i in range(H):
        for j in range(i, H):
            np.testing.assert_allclose(pop["hap_like"][0, 0, i], want_geno[g, 2], rtol=1e-12, atol=0)
            np.testing.assert_allclose(pop["hap_like"][0, 0, j], want_geno[g, 3], rtol=1e-12, atol=0)
            g += 1


def l3_pop_golden_cases(golden_dir):
    """Yields (batch, expected dict, use_em, (hla, flank)) for the windows of tests/golden/l3_pop_ref.npz (outputs of
    the reference's own Population class)."""
    import os
    import pickle
    import numpy as np
    g = np.load(os.path.join(golden_dir, "l3_pop_ref.npz"))
    for seed in range(int(g["n_cases"])):
        c, n_ind, mode, use_em, flat = l3_population_setup(seed)
        key = "p%d_" % seed
        off, hs = g[key + "hap_off"], g[key + "hap"]
        ref = {"hap_seq": [hs[off[k]:off[k + 1]].tobytes() for k in range(len(off) - 1)],
               "hap_start": int(g[key + "hap_start"]), "variants": pickle.loads(g[key + "variants"].tobytes())}
        b, phred = l3_population_batch(c, ref, flat)
        want = {k: g[key + k] for k in ("freq", "gl", "em", "gl_log_max", "gof", "call")}
        want["var_phred"] = phred
        yield b, want, use_em, mode


MANY_CASES = [(9001, 300), (9002, 300), (9003, 2000), (9004, 2000)]   # as tests/golden/make_golden.py


def l3_pop_many_cases(golden_dir):
    """Yields (batch, expected dict, use_em) for the many-sample windows of tests/golden/l3_pop_many_ref.npz (300 and 2000
    individuals through the reference's own Population class; BASELINE config 5's individual count)."""
    import os
    import pickle
    import numpy as np
    g = np.load(os.path.join(golden_dir, "l3_pop_many_ref.npz"))
    for k, (seed, n_ind) in enumerate(MANY_CASES):
        c = l3_window_case(seed, n_ind)
        key = "m%d_" % k
        off, hs = g[key + "hap_off"], g[key + "hap"]
        ref = {"hap_seq": [hs[off[j]:off[j + 1]].tobytes() for j in range(len(off) - 1)],
               "hap_start": int(g[key + "hap_start"]), "variants": pickle.loads(g[key + "variants"].tobytes())}
        b, phred = l3_population_batch(c, ref, True)
        want = {x: g[key + x] for x in ("freq", "gl", "em", "gl_log_max", "call")}
        want["stride"] = int(g[key + "stride"])
        want["var_phred"] = phred
        yield b, want, k % 2


def check_l3_pop_many(got, want):
    """Integer outputs equal the reference's; frequencies / posteriors to 1e-9 (they are equal unless exp / log of the two
    math libraries differ in a last bit)."""
    import numpy as np
    st = want["stride"]
    H, G = len(want["freq"]), want["gl"].shape[1]
    assert list(got["call"][0]) == list(want["call"]), "genotype calls"
    assert list(got["var_phred"][0, :len(want["var_phred"])]) == list(want["var_phred"]), "variant posteriors"
    np.testing.assert_allclose(got["freq"][0, :H], want["freq"], rtol=1e-9, atol=0, err_msg="freq")
    np.testing.assert_allclose(got["gl"][0, ::st, :G], want["gl"], rtol=1e-12, atol=0, err_msg="gl")
    np.testing.assert_allclose(got["em_post"][0, ::st, :G], want["em"], rtol=1e-9, atol=1e-300, err_msg="em")
    has = want["call"] >= 0
    np.testing.assert_allclose(got["gl_log_max"][0][has], want["gl_log_max"][has], rtol=1e-12, atol=0)


def check_l3_pop(got, want, rtol=1e-12):
    import numpy as np
    nI, G = want["gl"].shape
    H = len(want["freq"])
    np.testing.assert_allclose(got["gl"][0, :, :G], want["gl"], rtol=rtol, atol=0, err_msg="gl")
    np.testing.assert_allclose(got["gof"][0, :G, :], want["gof"], rtol=rtol, atol=0, err_msg="gof")
    np.testing.assert_allclose(got["freq"][0, :H], want["freq"], rtol=1e-9, atol=0, err_msg="freq")
    np.testing.assert_allclose(got["em_post"][0, :, :G], want["em"], rtol=1e-9, atol=1e-300, err_msg="em")
    has = want["call"] >= 0
    np.testing.assert_allclose(got["gl_log_max"][0][has], want["gl_log_max"][has], rtol=rtol, atol=0, err_msg="gl_log_max")
    assert list(got["call"][0]) == list(want["call"])
    assert list(got["var_phred"][0, :len(want["var_phred"])]) == list(want["var_phred"])


def n4_cases():
    """(batch, sites) pairs of the per-site genotype fixtures: 1, 3, 4 and 30 individuals (30 switches on the
    EM-frequency weighting, vcfutils.pyx:264-267), bi- and multi-allelic sites."""
    for n_ind, seed in ((1, 11), (3, 12), (30, 13), (4, 5)):
        b = edge_batch(seed=seed, n_windows=16, n_individuals=n_ind)
        yield b, sites_for_batch(b, seed=seed)


def check_n4(got, g, k, rtol=1e-12):
    """Site outputs against case k of tests/golden/n4_ref.npz (the reference's own function)."""
    import numpy as np
    have = g["n%d_have" % k].astype(bool)
    assert have.sum() > 20
    assert np.array_equal(got["phased"][have], g["n%d_phased" % k][have])
    P = g["n%d_lik" % k].shape[2]
    np.testing.assert_allclose(got["lik"][:, :, :P][have], g["n%d_lik" % k][have], rtol=rtol, atol=0)
    np.testing.assert_allclose(got["post"][have], g["n%d_post" % k][have], rtol=rtol, atol=0, equal_nan=True)
    np.testing.assert_allclose(got["gof"][have], g["n%d_gof" % k][have], rtol=rtol, atol=0)


# ---- N1: haplotype selection loop ------------------------------------------------------------------------------

def n1_window_case(seed, drop=0):
    """One window for getFilteredHaplotypes (variantFilter.pyx:377-506): 3-9 candidate variants (SNPs, 1-3 bp
    insertions / deletions, multi-allelic SNPs at one position whose unseen alleles give exactly tied scores, adjacent
    SNP + indel pairs, overlapping deletions that make some combinations invalid), nSupportingReads with ties, 1-3
    individuals (one may have no reads) with 8-90 good reads drawn from two true haplotypes, and the option values
    (small maxHaplotypes so that the heap overflows; coverage levels that switch sub-sampling on).  drop = 1 / 2 removes
    all reads / the first individual's reads afterwards (same variants and options)."""
    rng = random.Random(1000003 * seed + 17)
    genome = _rand_seq(rng, 2400)
    if seed % 4 == 0:
        genome = genome[:1215] + bytes([rng.choice(ACGT)]) * rng.randint(5, 12) + genome[1227:]
        genome = genome[:2400]
    ws = 1200 if seed % 11 else 120          # a window near the contig start: the left buffer is clamped
    we = ws + rng.randint(30, 70)
    max_read_len = rng.choice([100, 150]) if ws > 400 else 100
    n_var = rng.randint(3, 9)
    variants, pos = [], ws + rng.randint(0, 3)
    while len(variants) < n_var and pos < we - 3:
        kind = rng.random()
        if kind < 0.5:
            alts = [c for c in ACGT if c != genome[pos]]
            rng.shuffle(alts)
            for alt in sorted(alts[:rng.choice([1, 1, 1, 2, 3])]):
                variants.append((pos, genome[pos:pos + 1], bytes([alt])))
        elif kind < 0.7:
            variants.append((pos, b"", _rand_seq(rng, rng.randint(1, 3))))
        elif kind < 0.9:
            k = rng.randint(1, 4)
            variants.append((pos, genome[pos:pos + k], b""))
        else:
            k = rng.randint(2, 3)
            variants.append((pos, genome[pos:pos + k], _rand_seq(rng, k)))
        pos += rng.choice([0, 1, 1, 2, 3, 5, 8, 12])
    # the reference's window lists are sorted Variants (variant.pyx:304-315: position, type, nRemoved) without duplicates
    seen, uniq = set(), []
    for v in variants:
        if v not in seen:
            seen.add(v)
            uniq.append(v)

    def vtype(v):
        nr, na = len(v[1]), len(v[2])
        return (0 if na == 1 else 1) if nr == na else 2 if nr == 0 else 3 if na == 0 else 4
    uniq = uniq[:n_var]
    # variants that make getMutatedSequence run one base past the window end (the right buffer then repeats that base): a
    # deletion of the last window base, an insertion anchored on the window end.  (Further than one base the reference logs an
    # error, chaplotype.pyx:441-442 - in Python 3 that log call itself raises, so the fixture stays at one base; the library's
    # walk follows the same rule for any distance.)
    if seed % 7 == 3 and all(v[0] < we - 2 for v in uniq):
        uniq.append((we - 1, genome[we - 1:we], b""))
    elif seed % 7 == 5 and all(v[0] < we for v in uniq):
        uniq.append((we, b"", _rand_seq(rng, 2)))
    uniq.sort(key=lambda v: (v[0], vtype(v), len(v[1])))
    variants = [(p, r, a, rng.choice([1, 2, 2, 3, 5, 8, 20])) for (p, r, a) in uniq]
    flank = min(2 * max_read_len, 500)
    lo = max(0, ws - flank)

    def apply(idxs):
        out, cur = bytearray(), lo
        for i in idxs:
            p, rem, add, _ = variants[i]
            if p < cur:
                continue
            if len(rem) == len(add):
                out += genome[cur:p] + add
                cur = p + len(rem)
            elif not rem:
                out += genome[cur:p + 1] + add
                cur = p + 1
            else:
                out += genome[cur:p + 1]
                cur = p + 1 + len(rem)
        out += genome[cur:we + flank]
        return bytes(out)
    truth = []
    for _ in range(2):
        truth.append(apply(sorted(i for i in range(len(variants)) if rng.random() < 0.4)))
    n_ind = rng.choice([1, 1, 2, 3])
    per_ind = []
    for i in range(n_ind):
        if n_ind > 1 and i == 1 and rng.random() < 0.3:
            per_ind.append([])
            continue
        n = rng.choice([8, 15, 30, 60, 90])
        L0 = rng.choice([60, max_read_len])
        reads = []
        for _ in range(n):
            L = L0 if rng.random() < 0.8 else rng.choice([30, 45, 60])
            src = rng.choice(truth)
            a = max(0, ws - lo - L + 10)
            b = max(a, min(len(src) - L - 16, we - lo - 10))
            idx = rng.randint(a, b)
            seq = mutate(rng, src[idx:], L, n_rate=0.002)
            qual = bytes(rng.randint(2, 41) for _ in range(L))
            p = lo + idx + rng.choice([0, 0, 0, rng.randint(-4, 4)])
            reads.append((seq, qual, p, p + L, rng.choice([60, 60, 60, 40, 23, 0]), 512 if rng.random() < 0.03 else 0))
        reads.sort(key=lambda t: t[2])
        per_ind.append(reads)
    if drop == 1:      # no reads at all: every trial scores -1e20 and only the tuple order of the variant sets decides
        per_ind = [[] for _ in per_ind]
    elif drop == 2:    # the first individual has no reads
        per_ind = [[]] + per_ind[1:]
    opts = dict(max_haplotypes=rng.choice([4, 6, 9, 17, 50]), max_variants=8, filter_by_coverage=rng.choice([0, 0, 1]),
                coverage_sampling_level=rng.choice([3, 10, 30]))
    opts["original_max_haplotypes"] = opts["max_haplotypes"] if rng.random() < 0.7 else opts["max_haplotypes"] + rng.randint(1, 6)
    return dict(genome=genome, win_start=ws, win_end=we, variants=variants, per_ind=per_ind, max_read_len=max_read_len,
                opts=opts)


def n1_select_window(case, ref_seq, hap_start):
    """oracle.select_oracle.SelectWindow of an n1_window_case (ref_seq / hap_start as the reference built them)."""
    from oracle.select_oracle import SelectWindow
    good = [[Read(t[0], t[1], t[2], t[3], t[4], bool(t[5] & 512)) for t in ind] for ind in case["per_ind"]]
    return SelectWindow(ref_seq, case["win_start"], case["win_end"], hap_start, case["variants"], good)


def n1_batch(cases, ref_seqs, hap_starts):
    """(WindowBatch with ONE reference haplotype per window + the good reads, VariantSet) of n1_window_cases that
    share a number of individuals (windows with fewer individuals get empty read lists)."""
    from platypus_b200.batch import VariantSet
    n_ind = max(len(c["per_ind"]) for c in cases)
    wins = []
    for c, ref, hs in zip(cases, ref_seqs, hap_starts):
        per = [([Read(t[0], t[1], t[2], t[3], t[4], bool(t[5] & 512)) for t in ind], [], []) for ind in c["per_ind"]]
        per += [([], [], [])] * (n_ind - len(per))
        wins.append(Window(c["win_start"], c["win_end"], hs, [ref], per))
    return WindowBatch.from_windows(wins, n_ind, dedupe_reads=False), VariantSet.from_lists([c["variants"] for c in cases])


def hla_window_case(seed):
    """One window for getAllHLAHaplotypesInRegion (variantFilter.pyx:655-736): a few hundred known alleles (varSource
    FILE_VAR = 2; SNPs - up to three per position -, short insertions and deletions, pairs of insertions into a homopolymer
    that spell the SAME haplotype, and some read-derived variants, varSource 1, which get no haplotype), 1-2 individuals
    with 20-70 reads drawn from two true haplotypes.  Seeds divisible by 5 stay at or below the 150 haplotypes the
    function returns unfiltered."""
    rng = random.Random(7000003 * seed + 29)
    genome = bytearray(_rand_seq(rng, 3000))
    ws = 1300
    we = ws + rng.randint(220, 420)
    for _ in range(3):   # homopolymer runs: insertions of the run's base at two anchors give equal sequences
        a = rng.randint(ws + 5, we - 20)
        genome[a:a + rng.randint(4, 7)] = bytes([rng.choice(ACGT)]) * 7
    genome = bytes(genome[:3000])
    max_read_len = rng.choice([100, 150])
    target = rng.randint(60, 150) if seed % 5 == 0 else rng.randint(151, 330)
    var = set()
    while len(var) < target:
        p = rng.randint(ws, we - 4)
        kind = rng.random()
        if kind < 0.72:
            alt = rng.choice([c for c in ACGT if c != genome[p]])
            var.add((p, genome[p:p + 1], bytes([alt])))
        elif kind < 0.84:
            var.add((p, b"", _rand_seq(rng, rng.randint(1, 3))))
        elif kind < 0.92:
            var.add((p, b"", genome[p + 1:p + 2]))          # often inside a run: same sequence as its neighbour's insertion
        else:
            k = rng.randint(1, 4)
            var.add((p, genome[p:p + k], b""))

    def vtype(v):
        nr, na = len(v[1]), len(v[2])
        return (0 if na == 1 else 1) if nr == na else 2 if nr == 0 else 3 if na == 0 else 4
    # (a set of tuples holding bytes iterates in hash order, which changes from process to process: sort on everything)
    uniq = sorted(var, key=lambda v: (v[0], vtype(v), len(v[1]), v[1], v[2]))
    variants = [(p, r, a, rng.choice([1, 2, 3, 5, 8]), 1 if rng.random() < 0.06 else 2) for (p, r, a) in uniq]
    flank = min(2 * max_read_len, 500)
    lo = max(0, ws - flank)

    def apply(idxs):
        out, cur = bytearray(), lo
        for i in idxs:
            p, rem, add = variants[i][:3]
            if p < cur:
                continue
            if len(rem) == len(add):
                out += genome[cur:p] + add
                cur = p + len(rem)
            elif not rem:
                out += genome[cur:p + 1] + add
                cur = p + 1
            else:
                out += genome[cur:p + 1]
                cur = p + 1 + len(rem)
        out += genome[cur:we + flank]
        return bytes(out)
    truth = [apply(sorted(rng.sample(range(len(variants)), rng.randint(1, 4)))) for _ in range(2)]
    per_ind = []
    for i in range(rng.choice([1, 2])):
        reads = []
        for _ in range(rng.choice([20, 40, 70])):
            L = max_read_len if rng.random() < 0.8 else rng.choice([50, 75])
            src = rng.choice(truth)
            a = max(0, ws - lo - L + 10)
            b = max(a, min(len(src) - L - 16, we - lo - 10))
            idx = rng.randint(a, b)
            seq = mutate(rng, src[idx:], L, n_rate=0.001)
            qual = bytes(rng.randint(2, 41) for _ in range(L))
            p = lo + idx + rng.choice([0, 0, 0, rng.randint(-4, 4)])
            reads.append((seq, qual, p, p + L, rng.choice([60, 60, 60, 40, 23]), 512 if rng.random() < 0.02 else 0))
        reads.sort(key=lambda t: t[2])
        per_ind.append(reads)
    return dict(genome=genome, win_start=ws, win_end=we, variants=variants, per_ind=per_ind, max_read_len=max_read_len,
                opts=dict(original_max_haplotypes=rng.choice([20, 50, 50, 90]), coverage_sampling_level=rng.choice([5, 30])))


def hla_golden_cases(golden_dir):
    """The committed reference outputs for hla_window_case(seed) (tests/golden/make_n1_hla_fixture.py)."""
    z = np.load(os.path.join(golden_dir, "n1_hla_ref.npz"), allow_pickle=False)
    out = []
    for k, seed in enumerate(z["seeds"]):
        out.append(dict(seed=int(seed), haps=[int(i) for i in z["haps"][z["hap_off"][k]:z["hap_off"][k + 1]]],
                        hap_score=z["hap_score"][z["fv_off"][k]:z["fv_off"][k + 1]],
                        gt_score=z["gt_score"][z["fv_off"][k]:z["fv_off"][k + 1]],
                        ref_seq=z["ref_seq"][z["ref_off"][k]:z["ref_off"][k + 1]].tobytes(), hap_start=int(z["hap_start"][k])))
    return out


class MemFasta:
    """getSequence(refName, begin, end) over a bytes genome (half-open, clamped like fastafile.pyx:173-207)."""

    def __init__(self, genome):
        self.genome = genome

    def getSequence(self, name, begin, end):
        return self.genome[max(0, begin):min(len(self.genome), end)]


def hla_compat_inputs(case, engine):
    """(variants, refHaplotype, readBuffers, options) of an hla_window_case as platypus_b200.compat objects."""
    from platypus_b200 import compat
    opts = compat.Options(rlen=case["max_read_len"], HLATyping=0, originalMaxHaplotypes=case["opts"]["original_max_haplotypes"],
                          coverageSamplingLevel=case["opts"]["coverage_sampling_level"])
    fa = MemFasta(case["genome"])
    variants = [compat.Variant(b"chr", p, rem, add, n, None, src) for (p, rem, add, n, src) in case["variants"]]
    ref_hap = compat.Haplotype(b"chr", case["win_start"], case["win_end"], (), fa, case["max_read_len"], opts, engine)
    bufs = [compat.WindowReads(reads=ind) for ind in case["per_ind"]]
    return fa, variants, ref_hap, bufs, opts


def masks_of(sets):
    return [sum(1 << i for i in s) for s in sets]


def n1_golden_cases(golden_dir):
    """The committed reference outputs for n1_window_case(seed) (tests/golden/make_n1_fixture.py): per seed
    ref_seq, hap_start, selected masks, the sequences of the selected haplotypes, and the trial sets + scores of
    every round."""
    z = np.load(os.path.join(golden_dir, "n1_ref.npz"), allow_pickle=False)
    out = []
    for k, seed in enumerate(z["seeds"]):
        a, b = z["sel_off"][k], z["sel_off"][k + 1]
        t0, t1 = z["trial_off"][k], z["trial_off"][k + 1]
        seqs = [z["hap_seq"][z["hap_seq_off"][j]:z["hap_seq_off"][j + 1]].tobytes() for j in range(a, b)]
        out.append(dict(seed=int(seed), drop=int(z["drop"][k]), ref_seq=z["ref_seq"][z["ref_off"][k]:z["ref_off"][k + 1]].tobytes(),
                        hap_start=int(z["hap_start"][k]), sel_mask=[int(m) for m in z["sel_mask"][a:b]], hap_seqs=seqs,
                        trial_mask=[int(m) for m in z["trial_mask"][t0:t1]], trial_score=z["trial_score"][t0:t1],
                        hap_score=z["hap_score"][a:b], ref_hap_score=float(z["ref_hap_score"][k]),
                        opts={k2: int(z["opt_" + k2][k]) for k2 in ("max_haplotypes", "original_max_haplotypes", "max_variants",
                                                                   "filter_by_coverage", "coverage_sampling_level")}))
    return out


# ---- large differential test: adversarial windows ------------------------------------------------------------------

def adversarial_batch(seed, n_windows, n_haps=8, n_reads=60):
    """Windows built to reach every branch of mapAndAlignReadToHaplotype (calign.pyx:170-272) and every form of the band
    alignment through the WINDOW path: tandem repeats (tied vote maxima, many candidates), homopolymer runs of 30-60 bp
    (gap-open below gap-extend: the 6-op recurrence), haplotype N's (8-op form), IUPAC / lower-case bytes (byte-exact
    path), reads of 9-600 bp and one family above 2000 bp, reads unrelated to every haplotype, reads with N's, zeroed
    qualities, and BAM positions off by up to +-300.  Deterministic in (seed, n_windows)."""
    rng = random.Random(seed)
    rnd = np.random.default_rng(seed)

    def rseq(n):
        return ACGT_ARR[rnd.integers(0, 4, n)].tobytes()

    wins = []
    for w in range(n_windows):
        fam = w % 11
        long_reads = fam == 7
        huge = fam == 8 and w % 44 == 8
        Lmax = 2100 if huge else (rng.choice([300, 450, 600]) if long_reads else rng.choice([30, 100, 150, 150, 250]))
        hl = max(Lmax + rng.randint(40, 260), 130)
        base = bytearray(rseq(hl + 8))
        if fam in (0, 1):      # tandem repeat in the middle
            unit = rseq(rng.randint(1, 12))
            a = rng.randint(10, max(11, hl // 3))
            n_rep = rng.randint(20, max(21, (hl // 2) // len(unit)))
            rep = (unit * n_rep)[:hl - a - 10]
            base[a:a + len(rep)] = rep
        if fam in (2, 3):      # long homopolymer: gap-open drops below gap-extend inside runs >= 40
            a = rng.randint(10, hl - 80)
            base[a:a + rng.randint(30, 60)] = bytes([rng.choice(ACGT)]) * 60
            base = base[:hl + 8]
        if fam == 4:           # N's in the haplotypes
            for _ in range(rng.randint(1, 8)):
                p = rng.randrange(hl)
                base[p:p + rng.choice([1, 1, 2, 5])] = b"N" * 5
            base = base[:hl + 8]
        if fam == 5:           # bytes outside ACGTN
            for _ in range(rng.randint(1, 4)):
                base[rng.randrange(hl)] = rng.choice(b"RYKMSWacgtn")
        haps = [bytes(base[:hl])]
        c0 = hl // 2 - 25
        while len(haps) < n_haps:
            h = bytearray(base)
            for _ in range(rng.randint(1, 3)):
                p = c0 + rng.randrange(50)
                u = rng.random()
                if u < 0.6:
                    h[p] = rng.choice(ACGT)
                elif u < 0.8:
                    h[p:p] = rseq(rng.randint(1, 4))
                else:
                    del h[p:p + rng.randint(1, 4)]
            hb = bytes(h[:hl])
            if len(hb) == hl and hb not in haps:
                haps.append(hb)
            elif rng.random() < 0.05:
                haps.append(bytes(base[:hl - 1]) + bytes([rng.choice(ACGT)]))   # give up on distinct variants
        hs = 50000 + 3000 * w
        reads = []
        for _ in range(n_reads):
            if huge:
                L = rng.choice([2001, 2050, 2100])
            elif long_reads:
                L = rng.randint(250, Lmax)
            else:
                L = rng.choice([9, 10, 12, 20, 36, 50, 76, Lmax, Lmax, Lmax, rng.randint(9, Lmax)])
            L = min(L, hl - 16)
            src = rng.choice(haps)
            u = rng.random()
            idx = rng.randint(0, hl - L - 16)
            if u < 0.08 and L <= 450:
                seq = rseq(L)                                       # unrelated read (its score stays inside the int16 range)
            else:
                seq = mutate(rng, src[idx:] + rseq(8), L, sub=rng.choice([0.0, 0.01, 0.05]), ins=0.004, dele=0.004,
                             n_rate=0.003 if fam != 9 else 0.02)
            q = rnd.integers(2, 42, L).astype(np.uint8)
            if rng.random() < 0.1:
                q[rnd.integers(0, L, max(1, L // 10))] = 0
            jit = rng.choice([0, 0, 0, 0, rng.randint(-8, 8), rng.randint(-300, 300)])
            pos = hs + idx + jit
            reads.append(Read(seq, q.tobytes(), pos, pos + L, rng.choice([60, 60, 60, 37, 12, 0]), rng.random() < 0.02))
        ng = rng.randint(n_reads // 2, n_reads)
        nb = rng.randint(0, n_reads - ng)
        if w % 5 == 4:       # a 50 bp interval: the overlap < 7 rule short-circuits many good / bad reads
            ws, we = hs + c0, hs + c0 + 50
        else:                # the interval spans the haplotype: nearly every read is scored
            ws, we = hs + 5, hs + hl - 5
        wins.append(Window(ws, we, hs, haps, [(reads[:ng], reads[ng:ng + nb], reads[ng + nb:])]))
    return WindowBatch.from_windows(wins, 1)


ACGT_ARR = np.frombuffer(ACGT, np.uint8)
