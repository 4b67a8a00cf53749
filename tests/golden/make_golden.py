"""
Generates the golden fixtures in this directory.  Run in the BUILD container, where
/root/reference exists; the GPU box only reads the committed .npz files.

  align_ref.npz   inputs + scores from the UNMODIFIED reference kernel
                  (src/c/align.c fastAlignmentRoutine via oracle/_ref/libalign_ref.so)
  calign_ref.npz  inputs + scores from the reference's src/cython/calign.pyx
                  mapAndAlignReadToHaplotype (oracle/_ref/calign*.so), gap-open tables from the
                  oracle's restatement of chaplotype.pyx:552-590
  align_tb_ref.npz   inputs + traceback rows / firstpos of fastAlignmentRoutine and
                  calculateFlankScore values (src/c/align.c:523-644) from the same library
  calign_modes_ref.npz  mapAndAlignReadToHaplotype with doCalculateFlankScore = 1 and with
                  HLA-style clipped reads whose hashes stay those of the unclipped read
                  (chaplotype.pyx:637-655), from the reference's calign.pyx
  l3_ref.npz      per-read log-likelihoods (Haplotype.alignReads) and genotype log-likelihoods / GOF / hapLike
                  (DiploidGenotype.calculateDataLikelihood) from the reference's own chaplotype.pyx / cgenotype.pyx
                  (oracle/_ref/l3_ref_wrap*.so) for the windows of tests/cases.l3_window_case, in default, HLA and
                  flank mode; haplotype sequences as the reference's Haplotype constructor built them
  l3_pop_ref.npz  the reference's own Population class (cpopulation.pyx setup() + call()): rescaled genotype
                  likelihoods, maxLogLikelihoods, GOF, EM haplotype frequencies, EM genotype posteriors, genotype
                  calls and calculatePosterior values for multi-individual windows (tests/cases.l3_population_setup)
  n4_ref.npz      computeGenotypeCallAndLikelihoods of the reference (vcfutils.pyx:163-334, excerpted at build time into
                  oracle/_ref/n4_ref*.so) for every (site, individual) of tests/cases.n4_cases: phased indices, marginal
                  likelihoods, genotype / non-ref / ref posteriors, best GOF
  window_modes_restated.npz  the edge batch under --calculateFlankScore=1 / --HLATyping=1 from the
                  oracle (restated above the integer score)
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


def make_align_tb(n=400, seed=21):
    assert O.ref_align_lib() is not None
    rng = random.Random(seed)
    haps, gos, reads, quals = [], [], [], []
    scores, a1s, a2s, fps, flanks, hflank = [], [], [], [], [], []
    for i in range(n):
        hap, go, read, qual = cases.random_alignment_case(rng, i)
        s, a1, a2, fp = O.ref_fast_align_tb(hap, read, qual, go)
        fl = rng.randint(1, max(1, len(hap) // 2))
        f = O.ref_flank_score(len(hap), fl, qual, go, fp, a1, a2)
        haps.append(hap); gos.append(go[:len(hap)]); reads.append(read); quals.append(qual)
        scores.append(s); a1s.append(a1); a2s.append(a2); fps.append(fp); flanks.append(f); hflank.append(fl)
    ho, hs = pack(haps)
    _, gs = pack(gos)
    ro, rs = pack(reads)
    _, qs = pack(quals)
    ao, a1 = pack(a1s)
    _, a2 = pack(a2s)
    np.savez_compressed(os.path.join(HERE, "align_tb_ref.npz"), hap_off=ho, hap=hs, gap_open=gs, read_off=ro, read=rs,
                        qual=qs, score=np.array(scores, np.int32), aln_off=ao, aln1=a1, aln2=a2,
                        firstpos=np.array(fps, np.int32), hap_flank=np.array(hflank, np.int32),
                        flank_score=np.array(flanks, np.int32))
    print("align_tb_ref.npz:", n, "cases,", sum(1 for f in flanks if f), "with a nonzero flank score")


def make_calign_modes(n=400, seed=22):
    cw = O.ref_calign()
    assert cw is not None
    rng = random.Random(seed)
    haps, reads, quals, hreads = [], [], [], []
    rstart, hstart, hflank, doflank, scores = [], [], [], [], []
    for i in range(n):
        hap, read, qual, read_start, hap_start = cases.random_mapping_case(rng, i)
        go = O.gap_open(hap)
        fl = rng.randint(1, max(1, len(hap) // 2))
        do = 0 if i % 4 == 3 else 1
        hread = read
        if i % 2 == 1 and len(read) > 30:   # HLA-style clip; votes still come from the whole read
            off1 = rng.choice([0, 1, 3, 10]); off2 = rng.choice([0, 2, 7])
            read, qual, read_start = read[off1:len(read) - off2], qual[off1:len(qual) - off2], read_start + off1
        s = cw.map_and_align(read, qual, read_start, hap_start, hap, go, 3, 2, fl, do, 0, hread)
        haps.append(hap); reads.append(read); quals.append(qual); hreads.append(hread)
        rstart.append(read_start); hstart.append(hap_start); hflank.append(fl); doflank.append(do); scores.append(s)
    ho, hs = pack(haps)
    ro, rs = pack(reads)
    _, qs = pack(quals)
    uo, us = pack(hreads)
    np.savez_compressed(os.path.join(HERE, "calign_modes_ref.npz"), hap_off=ho, hap=hs, read_off=ro, read=rs, qual=qs,
                        hash_read_off=uo, hash_read=us, read_start=np.array(rstart, np.int32),
                        hap_start=np.array(hstart, np.int32), hap_flank=np.array(hflank, np.int32),
                        do_flank=np.array(doflank, np.int32), score=np.array(scores, np.int32))
    print("calign_modes_ref.npz:", n, "cases")


def make_window_modes():
    from platypus_b200 import _abi
    out = {}
    for name, kw in (("flank", dict(calc_flank_score=1)), ("hla", dict(use_mapq_cap=1)),
                     ("both", dict(calc_flank_score=1, use_mapq_cap=1))):
        batch = cases.edge_batch(seed=5, overhang=True)
        opt = _abi.PlbOptions.default()
        for k, v in kw.items():
            setattr(opt, k, v)
        arrs, ll, sc, st = O.population_run(batch, opt)
        out[name + "_ll"] = ll
        out[name + "_score"] = sc
        for k in ("gl", "freq", "var_phred", "call"):
            out[name + "_" + k] = arrs[k]
    np.savez_compressed(os.path.join(HERE, "window_modes_restated.npz"), **out)
    print("window_modes_restated.npz written")


def make_l3(n=60):
    W = O.ref_l3()
    assert W is not None, "reference chaplotype.pyx / cgenotype.pyx not built (need /root/reference + Cython)"
    out = {}
    modes = [(0, 0), (1, 0), (0, 1), (1, 1)]
    n_ll = 0
    for seed in range(n):
        c = cases.l3_window_case(seed)
        for hla, flank in modes:
            r = W.window_likelihoods(c["genome"], c["win_start"], c["win_end"], c["hap_variants"], c["good"], c["bad"],
                                     c["broken"], c["max_read_len"], hla, flank)
            key = "c%d_m%d%d_" % (seed, hla, flank)
            if (hla, flank) == (0, 0):
                ho, hs = pack(r["hap_seq"])
                out["c%d_hap_off" % seed] = ho
                out["c%d_hap" % seed] = hs
                out["c%d_hap_start" % seed] = np.int32(r["hap_start"])
            out[key + "ll"] = np.array(r["ll"], np.float64)
            out[key + "geno"] = np.array([g[2:] for g in r["genotypes"]], np.float64)   # logL, gof, hap1Like, hap2Like
            n_ll += out[key + "ll"].size
    np.savez_compressed(os.path.join(HERE, "l3_ref.npz"), n_cases=np.int32(n), **out)
    print("l3_ref.npz:", n, "windows x", len(modes), "modes,", n_ll, "log-likelihoods")


def make_l3_pop(n=48):
    import pickle
    W = O.ref_l3()
    assert W is not None
    out = {}
    for seed in range(n):
        c, n_ind, (hla, flank), use_em, flat = cases.l3_population_setup(seed)
        r = W.population(c["genome"], c["win_start"], c["win_end"], c["hap_variants"], c["per_ind"], c["max_read_len"], hla,
                         flank, use_em)
        key = "p%d_" % seed
        ho, hs = pack(r["hap_seq"])
        out[key + "hap_off"], out[key + "hap"], out[key + "hap_start"] = ho, hs, np.int32(r["hap_start"])
        for k in ("freq", "gl", "em", "gl_log_max", "gof"):
            out[key + k] = np.array(r[k], np.float64)
        out[key + "call"] = np.array(r["call"], np.int32)
        # variants: (refPos, removed, added, phred under the flat prior, prior or None, phred under it or None, holders)
        out[key + "variants"] = np.frombuffer(pickle.dumps(r["variants"], protocol=2), np.uint8)
    np.savez_compressed(os.path.join(HERE, "l3_pop_ref.npz"), n_cases=np.int32(n), **out)
    print("l3_pop_ref.npz:", n, "windows")


MANY_CASES = [(9001, 300), (9002, 300), (9003, 2000), (9004, 2000)]   # (seed, individuals) of l3_pop_many_ref.npz


def make_l3_pop_many():
    """Many-sample windows (BASELINE config 5: 2000 individuals) through the reference's own Population class: the EM's
    sums over individuals are long floating-point chains whose order matters for the last bits (cpopulation.pyx:384-457),
    so calls, posteriors and frequencies are pinned at that size too.  Inputs are regenerated from the seed
    (tests/cases.l3_window_case); stored: frequencies, calls, max log-likelihoods, variant posteriors, and every
    `stride`-th row of the genotype likelihoods / EM posteriors."""
    import pickle
    W = O.ref_l3()
    assert W is not None
    out = {}
    for k, (seed, n_ind) in enumerate(MANY_CASES):
        c = cases.l3_window_case(seed, n_ind)
        r = W.population(c["genome"], c["win_start"], c["win_end"], c["hap_variants"], c["per_ind"], c["max_read_len"], 0, 0, k % 2)
        key = "m%d_" % k
        ho, hs = pack(r["hap_seq"])
        out[key + "hap_off"], out[key + "hap"], out[key + "hap_start"] = ho, hs, np.int32(r["hap_start"])
        stride = 1 if n_ind <= 300 else 25
        out[key + "stride"] = np.int32(stride)
        out[key + "freq"] = np.array(r["freq"], np.float64)
        out[key + "gl_log_max"] = np.array(r["gl_log_max"], np.float64)
        out[key + "gl"] = np.array(r["gl"], np.float64)[::stride]
        out[key + "em"] = np.array(r["em"], np.float64)[::stride]
        out[key + "call"] = np.array(r["call"], np.int32)
        out[key + "variants"] = np.frombuffer(pickle.dumps(r["variants"], protocol=2), np.uint8)
        print("many-sample case %d: %d individuals, %d haplotypes, freq %s" % (k, n_ind, len(r["freq"]), np.round(r["freq"], 4)))
    np.savez_compressed(os.path.join(HERE, "l3_pop_many_ref.npz"), n_cases=np.int32(len(MANY_CASES)), **out)
    print("l3_pop_many_ref.npz:", os.path.getsize(os.path.join(HERE, "l3_pop_many_ref.npz")), "bytes")


def make_n4():
    assert O.ref_n4() is not None
    out = {}
    tot = 0
    for k, (b, sites) in enumerate(cases.n4_cases()):
        pop, _, _, _ = O.population_run(b)
        ref = O.ref_site_genotypes(b, pop, sites)
        S, nI, P = sites.n_sites, b.n_individuals, sites.max_pairs()
        phased = np.full((S, nI, 2), -1, np.int32)
        lik = np.zeros((S, nI, P))
        post = np.zeros((S, nI, 3))
        gof = np.zeros((S, nI))
        have = np.zeros((S, nI), np.uint8)
        for (s_, i), (p1, p2, liks, gp, npo, rp, gf) in ref.items():
            phased[s_, i] = (p1, p2)
            lik[s_, i, :len(liks)] = liks
            post[s_, i] = (gp, npo, rp)
            gof[s_, i] = gf
            have[s_, i] = 1
        tot += int(have.sum())
        out.update({"n%d_phased" % k: phased, "n%d_lik" % k: lik, "n%d_post" % k: post, "n%d_gof" % k: gof, "n%d_have" % k: have})
    np.savez_compressed(os.path.join(HERE, "n4_ref.npz"), **out)
    print("n4_ref.npz:", tot, "(site, individual) cases")


def make_window():
    batch = cases.edge_batch(seed=5)
    arrs, ll, sc, st = O.population_run(batch)
    out = {k: v for k, v in arrs.items() if k != "max_haps"}
    np.savez_compressed(os.path.join(HERE, "window_restated.npz"), ll=ll, score=sc, **out)
    print("window_restated.npz:", batch.n_windows, "windows", st)


if __name__ == "__main__":
    make_align()
    make_calign()
    make_align_tb()
    make_calign_modes()
    make_l3()
    make_l3_pop()
    make_l3_pop_many()
    make_n4()
    make_window()
    make_window_modes()
