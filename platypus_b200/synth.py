"""
"synth-v1" synthetic windows: the workload BASELINE.json's configs are quoted on.

Counter-based RNG (numpy Philox) keyed by (seed, chunk of 256 windows), so window w has the
same bytes no matter how the job is sharded or how many windows are generated around it.

Per window: a reference segment of hapLen+16 iid ACGT (15 % chance of one homopolymer run of
4-12 in the middle third); haplotype 0 = reference, haplotypes 1..H-1 carry 1-3 variants
(SNP 70 % / 1-3 bp insertion 15 % / 1-3 bp deletion 15 %) at distinct positions of the central
50 bp and are re-cut to hapLen, never duplicating another haplotype (the reference dedupes,
src/cython/variantcaller.pyx:325-390).  Reads are drawn from one of two "true" haplotypes:
start uniform in [0, hapLen-L-16], qualities 90 % U[25,40] / 10 % U[2,24], substitutions with
probability 10^(-q/10), 0.2 %/base 1-bp indel errors, mapq 85 % 60 / 10 % U[20,59] / 5 %
U[0,19], read.pos = hapStart + idx + jitter (0 w.p. 0.9 else U[-5,5]).
"""
import numpy as np

from .batch import WindowBatch

SEED = 20261017
CHUNK = 256
_ACGT = np.frombuffer(b"ACGT", np.uint8)


def _rng(seed, chunk_id):
    return np.random.Generator(np.random.Philox(key=[seed, chunk_id]))


def _make_haps(rng, H, hap_len):
    """Returns (list of H uint8 arrays of length hap_len, list of variant sets, variant list)."""
    ref = _ACGT[rng.integers(0, 4, hap_len + 16)]
    if rng.random() < 0.15:
        run = int(rng.integers(4, 13))
        p = int(rng.integers(hap_len // 3, 2 * hap_len // 3 - run))
        ref[p:p + run] = ref[p]
    c0 = (hap_len - 50) // 2
    haps = [ref[:hap_len].copy()]
    seen = {haps[0].tobytes()}
    variants = []      # (pos, kind, payload)
    var_ix = {}
    hap_vars = [set()]
    tries = 0
    while len(haps) < H:
        tries += 1
        nv = int(rng.integers(1, 4))
        pos = sorted(int(x) for x in rng.choice(50, nv, replace=False) + c0)
        edits = []
        for p in pos:
            u = rng.random()
            if u < 0.70:
                alt = _ACGT[(int(np.searchsorted(_ACGT, ref[p])) + int(rng.integers(1, 4))) % 4]
                edits.append((p, "S", bytes([alt])))
            elif u < 0.85:
                edits.append((p, "I", _ACGT[rng.integers(0, 4, int(rng.integers(1, 4)))].tobytes()))
            else:
                edits.append((p, "D", int(rng.integers(1, 4))))
        out, cur = [], 0
        for p, k, pl in edits:
            if p < cur:
                continue
            out.append(ref[cur:p])
            if k == "S":
                out.append(np.frombuffer(pl, np.uint8))
                cur = p + 1
            elif k == "I":
                out.append(ref[p:p + 1])
                out.append(np.frombuffer(pl, np.uint8))
                cur = p + 1
            else:
                cur = p + pl
        out.append(ref[cur:])
        seq = np.concatenate(out)
        if len(seq) < hap_len:
            continue
        seq = seq[:hap_len].copy()
        key = seq.tobytes()
        if key in seen and tries < 200:
            continue
        seen.add(key)
        vs = set()
        for e in edits:
            if e not in var_ix:
                var_ix[e] = len(variants)
                variants.append(e)
            vs.add(var_ix[e])
        haps.append(seq)
        hap_vars.append(vs)
    return haps, hap_vars, variants


def make_batch(n_windows, n_haps=8, n_reads=64, read_len=150, hap_len=250, n_individuals=1, seed=SEED,
               window_offset=0, read_len_range=None, hap_len_range=None, with_variants=True, pos_jitter=0.1):
    """Build a WindowBatch of synth-v1 windows [window_offset, window_offset+n_windows).

    read_len_range=(lo,hi): per-read length U{lo..hi} (config 3); hap_len_range=(lo,hi): per-window
    haplotype length U[max(lo, Lmax+16), hi].  n_reads is per individual.
    """
    W = n_windows
    nI = n_individuals
    win_hap_off = np.arange(W + 1, dtype=np.int32) * n_haps
    hap_lens = np.zeros(W * n_haps, np.int64)
    hap_chunks = []
    R = n_reads * nI
    read_lens_all, read_seq_chunks, read_qual_chunks = [], [], []
    read_pos = np.zeros(W * R, np.int32)
    read_mapq = np.zeros(W * R, np.uint8)
    hap_start = np.zeros(W, np.int32)
    win_start = np.zeros(W, np.int32)
    win_end = np.zeros(W, np.int32)
    masks = np.zeros(W * n_haps, np.uint64)
    n_var = np.zeros(W, np.int32)
    priors = []
    Lmax = read_len_range[1] if read_len_range else read_len

    first_chunk = window_offset // CHUNK
    last_chunk = (window_offset + W - 1) // CHUNK if W else first_chunk - 1
    for ck in range(first_chunk, last_chunk + 1):
        rng = _rng(seed, ck)
        for wg in range(ck * CHUNK, (ck + 1) * CHUNK):
            # every window of the chunk is generated (cheaply skipped ones still advance the RNG)
            hl = hap_len if not hap_len_range else int(rng.integers(max(hap_len_range[0], Lmax + 16), hap_len_range[1] + 1))
            haps, hap_vars, variants = _make_haps(rng, n_haps, hl)
            if read_len_range:
                L = rng.integers(read_len_range[0], read_len_range[1] + 1, R)
            else:
                L = np.full(R, read_len, np.int64)
            g1, g2 = int(rng.integers(0, n_haps)), int(rng.integers(0, n_haps))
            src_hap = np.where(rng.random(R) < 0.5, g1, g2)
            idx = (rng.random(R) * (hl - L - 16 + 1)).astype(np.int64)
            Lm = int(L.max())
            q = np.where(rng.random((R, Lm)) < 0.9, rng.integers(25, 41, (R, Lm)), rng.integers(2, 25, (R, Lm))).astype(np.uint8)
            ev = rng.random((R, Lm))
            ins = ev < 0.001
            dele = (ev >= 0.001) & (ev < 0.002)
            adv = np.ones((R, Lm), np.int64) - ins + dele
            srcpos = idx[:, None] + np.cumsum(adv, axis=1) - adv
            srcpos = np.minimum(srcpos, hl - 1)
            hap2d = np.stack(haps)
            bases = hap2d[src_hap[:, None], srcpos]
            rnd_base = _ACGT[rng.integers(0, 4, (R, Lm))]
            bases = np.where(ins, rnd_base, bases)
            sub = rng.random((R, Lm)) < np.power(10.0, -q.astype(np.float64) / 10.0)
            code = np.searchsorted(_ACGT, bases)
            alt = _ACGT[(code + rng.integers(1, 4, (R, Lm))) % 4]
            bases = np.where(sub, alt, bases).astype(np.uint8)
            u = rng.random(R)
            mq = np.where(u < 0.85, 60, np.where(u < 0.95, rng.integers(20, 60, R), rng.integers(0, 20, R)))
            jit = np.where(rng.random(R) < (1.0 - pos_jitter), 0, rng.integers(-5, 6, R))
            if wg < window_offset or wg >= window_offset + W:
                continue
            w = wg - window_offset
            hs = 100000 + wg * 1000
            hap_start[w] = hs
            win_start[w] = hs + (hl - 50) // 2
            win_end[w] = win_start[w] + 50
            hap_lens[w * n_haps:(w + 1) * n_haps] = hl
            hap_chunks.append(hap2d.reshape(-1))
            valid = np.arange(Lm)[None, :] < L[:, None]
            read_seq_chunks.append(bases[valid])
            read_qual_chunks.append(q[valid])
            read_lens_all.append(L)
            read_pos[w * R:(w + 1) * R] = hs + idx + jit
            read_mapq[w * R:(w + 1) * R] = mq
            n_var[w] = min(len(variants), 64)
            for h in range(n_haps):
                m = 0
                for v in hap_vars[h]:
                    if v < 64:
                        m |= (1 << v)
                masks[w * n_haps + h] = m
            priors.append([1e-3 if v[1] == "S" else 1e-4 for v in variants[:64]])

    read_lens = np.concatenate(read_lens_all) if read_lens_all else np.zeros(0, np.int64)
    read_seq_off = np.zeros(W * R + 1, np.int64)
    np.cumsum(read_lens, out=read_seq_off[1:])
    hap_seq_off = np.zeros(W * n_haps + 1, np.int64)
    np.cumsum(hap_lens, out=hap_seq_off[1:])
    b = WindowBatch(
        n_windows=W, n_individuals=nI, win_hap_off=win_hap_off, win_start=win_start, win_end=win_end,
        hap_start=hap_start, hap_seq_off=hap_seq_off,
        hap_seq=np.concatenate(hap_chunks) if hap_chunks else np.zeros(0, np.uint8),
        wi_slot_off=np.arange(W * nI + 1, dtype=np.int64) * n_reads,
        wi_n_good=np.full(W * nI, n_reads, np.int32), wi_n_bad=np.zeros(W * nI, np.int32),
        slot_read=np.arange(W * R, dtype=np.int32),
        read_seq_off=read_seq_off,
        read_seq=np.concatenate(read_seq_chunks) if read_seq_chunks else np.zeros(0, np.uint8),
        read_qual=np.concatenate(read_qual_chunks) if read_qual_chunks else np.zeros(0, np.uint8),
        read_pos=read_pos, read_end=(read_pos + read_lens).astype(np.int32), read_mapq=read_mapq,
        read_qcfail=np.zeros(W * R, np.uint8),
    )
    if with_variants and W:
        mv = max(1, int(n_var.max()))
        pri = np.zeros((W, mv), np.float64)
        for w, p in enumerate(priors):
            pri[w, :len(p)] = p
        b.max_variants = mv
        b.win_n_var = n_var
        b.hap_var_mask = masks
        b.var_prior = pri
    return b


def scored_slots(batch: WindowBatch) -> np.ndarray:
    """Boolean per slot: True where the reference scores the read, False where Haplotype.alignReads
    short-circuits it to LL = 0 (QC fail or < 7 bp overlap with the window, good and bad reads only;
    reference: src/cython/chaplotype.pyx:343-361)."""
    nI = batch.n_individuals
    n_slots = batch.n_slots
    wi_of_slot = np.repeat(np.arange(batch.n_windows * nI), np.diff(batch.wi_slot_off))
    t = np.arange(n_slots) - batch.wi_slot_off[wi_of_slot]
    checked = t < (batch.wi_n_good + batch.wi_n_bad)[wi_of_slot]
    w = wi_of_slot // nI
    r = batch.slot_read
    s = np.maximum(batch.win_start[w], batch.read_pos[r])
    e = np.minimum(batch.win_end[w], batch.read_end[r])
    overlap = np.where(e > s, e - s, -1)
    skip = checked & ((batch.read_qcfail[r] != 0) | (overlap < 7))
    return ~skip


def algorithmic_cells(batch: WindowBatch) -> int:
    """16 * readLen per scored (read, haplotype) pair; pairs short-circuited by the
    QC-fail / overlap rule count 0."""
    nI = batch.n_individuals
    H = np.repeat(batch.haps_per_window().astype(np.int64), nI)
    lens = np.diff(batch.read_seq_off)
    slot_len = np.where(scored_slots(batch), lens[batch.slot_read], 0).astype(np.int64)
    cs = np.concatenate([[0], np.cumsum(slot_len)])
    per_wi = cs[batch.wi_slot_off[1:]] - cs[batch.wi_slot_off[:-1]]
    return int((per_wi * H).sum()) * 16


def algorithmic_bytes(batch: WindowBatch) -> int:
    """Algorithmic HBM bytes of one pass: per read ceil(L/4)+L+8, per haplotype
    ceil(hapLen/4)+8, outputs 8*H*T (per-read LL) + 8*nInd*G (GL)."""
    lens = np.diff(batch.read_seq_off)
    slot_len = lens[batch.slot_read].astype(np.int64)
    rd = int(((slot_len + 3) // 4 + slot_len + 8).sum())
    hl = np.diff(batch.hap_seq_off).astype(np.int64)
    hp = int(((hl + 3) // 4 + 8).sum())
    H = batch.haps_per_window().astype(np.int64)
    T = np.diff(batch.wi_slot_off).astype(np.int64)
    out = int((np.repeat(H, batch.n_individuals) * T).sum()) * 8
    out += int((H * (H + 1) // 2).sum()) * 8 * batch.n_individuals
    return rd + hp + out


def _make_range(args):
    lo, hi, kw = args
    return make_batch(hi - lo, window_offset=lo, **kw)


def make_batch_parallel(n_windows, window_offset=0, n_procs=None, **kw):
    """make_batch() over a process pool (chunk-aligned ranges, so the bytes are identical)."""
    import multiprocessing as mp
    import os
    from .batch import concat_batches
    n_procs = n_procs or min(os.cpu_count() or 1, 32)
    if n_procs <= 1 or n_windows <= 2 * CHUNK:
        return make_batch(n_windows, window_offset=window_offset, **kw)
    lo, hi = window_offset, window_offset + n_windows
    cuts = sorted(set([lo, hi] + list(range((lo // CHUNK + 1) * CHUNK, hi, CHUNK))))
    step = max(1, (len(cuts) - 1 + n_procs * 2 - 1) // (n_procs * 2))
    cuts = cuts[::step] + ([hi] if cuts[::step][-1] != hi else [])
    jobs = [(a, b, kw) for a, b in zip(cuts[:-1], cuts[1:])]
    with mp.get_context("fork").Pool(n_procs) as pool:
        parts = pool.map(_make_range, jobs)
    return concat_batches(parts)


# ---- N1 workload: windows for the haplotype selection loop ("synth-select-v1") ---------------------------------------

def make_select_batch(n_windows, n_vars=8, n_reads=64, read_len=150, hap_len=250, n_individuals=1, seed=SEED + 1,
                      window_offset=0):
    """Windows for getFilteredHaplotypes (src/cython/variantFilter.pyx:377-506), one RNG stream per window.

    Per window: a reference segment of hap_len iid ACGT (15 % with a homopolymer run), interval = the central 50 bp,
    n_vars candidate variants (SNP 70 % / 1-3 bp insertion 15 % / 1-3 bp deletion 15 %) at distinct positions of the
    interval at least 4 bp apart (so every combination is a valid haplotype), nSupportingReads U[1,20]; two true
    haplotypes carry each variant with probability 0.35; reads as in synth-v1 (same qualities, error rates, mapq mix
    and position jitter), all good, sorted by position as a read buffer holds them.
    Returns (WindowBatch with ONE reference haplotype per window, VariantSet)."""
    from .batch import VariantSet
    W, nI = n_windows, n_individuals
    R = n_reads * nI
    assert n_vars <= 12
    c0 = (hap_len - 50) // 2
    ref_chunks, read_seq_chunks, read_qual_chunks, per_window_vars = [], [], [], []
    read_pos = np.zeros(W * R, np.int32)
    read_mapq = np.zeros(W * R, np.uint8)
    hap_start = np.zeros(W, np.int32)
    win_start = np.zeros(W, np.int32)
    for w in range(W):
        wg = window_offset + w
        rng = np.random.Generator(np.random.Philox(key=[seed, wg]))
        ref = _ACGT[rng.integers(0, 4, hap_len)]
        if rng.random() < 0.15:
            run = int(rng.integers(4, 13))
            p = int(rng.integers(hap_len // 3, 2 * hap_len // 3 - run))
            ref[p:p + run] = ref[p]
        hs = 100000 + wg * 1000
        pos = np.sort(rng.choice(12, n_vars, replace=False)) * 4 + c0 + 1
        variants = []
        for p in pos:
            p = int(p)
            u = rng.random()
            if u < 0.70:
                alt = _ACGT[(int(np.searchsorted(_ACGT, ref[p])) + int(rng.integers(1, 4))) % 4]
                variants.append((hs + p, 1, bytes([alt]), int(rng.integers(1, 21))))
            elif u < 0.85:
                variants.append((hs + p, 0, _ACGT[rng.integers(0, 4, int(rng.integers(1, 4)))].tobytes(), int(rng.integers(1, 21))))
            else:
                variants.append((hs + p, int(rng.integers(1, 4)), b"", int(rng.integers(1, 21))))
        truth = []
        for _ in range(2):
            out, cur = [], 0
            for (gp, nrem, add, _n) in variants:
                if rng.random() >= 0.35:
                    continue
                p = gp - hs
                if nrem == len(add):
                    out += [ref[cur:p], np.frombuffer(add, np.uint8)]
                    cur = p + nrem
                elif nrem == 0:
                    out += [ref[cur:p + 1], np.frombuffer(add, np.uint8)]
                    cur = p + 1
                else:
                    out.append(ref[cur:p + 1])
                    cur = p + 1 + nrem
            out.append(ref[cur:])
            truth.append(np.concatenate(out))
        tl = min(len(t) for t in truth)
        hap2d = np.stack([t[:tl] for t in truth])
        L = read_len
        src = rng.integers(0, 2, R)
        idx = (rng.random(R) * (tl - L - 16 + 1)).astype(np.int64)
        q = np.where(rng.random((R, L)) < 0.9, rng.integers(25, 41, (R, L)), rng.integers(2, 25, (R, L))).astype(np.uint8)
        ev = rng.random((R, L))
        ins = ev < 0.001
        dele = (ev >= 0.001) & (ev < 0.002)
        adv = np.ones((R, L), np.int64) - ins + dele
        srcpos = np.minimum(idx[:, None] + np.cumsum(adv, axis=1) - adv, tl - 1)
        bases = hap2d[src[:, None], srcpos]
        bases = np.where(ins, _ACGT[rng.integers(0, 4, (R, L))], bases)
        sub = rng.random((R, L)) < np.power(10.0, -q.astype(np.float64) / 10.0)
        alt = _ACGT[(np.searchsorted(_ACGT, bases) + rng.integers(1, 4, (R, L))) % 4]
        bases = np.where(sub, alt, bases).astype(np.uint8)
        u = rng.random(R)
        mq = np.where(u < 0.85, 60, np.where(u < 0.95, rng.integers(20, 60, R), rng.integers(0, 20, R)))
        jit = np.where(rng.random(R) < 0.9, 0, rng.integers(-5, 6, R))
        rp = hs + idx + jit
        # per individual, reads sorted by position (stable)
        order = np.concatenate([i * n_reads + np.argsort(rp[i * n_reads:(i + 1) * n_reads], kind="stable") for i in range(nI)])
        read_seq_chunks.append(bases[order].reshape(-1))
        read_qual_chunks.append(q[order].reshape(-1))
        read_pos[w * R:(w + 1) * R] = rp[order]
        read_mapq[w * R:(w + 1) * R] = mq[order]
        hap_start[w] = hs
        win_start[w] = hs + c0
        ref_chunks.append(ref)
        per_window_vars.append(variants)
    b = WindowBatch(
        n_windows=W, n_individuals=nI, win_hap_off=np.arange(W + 1, dtype=np.int32), win_start=win_start,
        win_end=(win_start + 50).astype(np.int32), hap_start=hap_start,
        hap_seq_off=np.arange(W + 1, dtype=np.int64) * hap_len,
        hap_seq=np.concatenate(ref_chunks) if W else np.zeros(0, np.uint8),
        wi_slot_off=np.arange(W * nI + 1, dtype=np.int64) * n_reads,
        wi_n_good=np.full(W * nI, n_reads, np.int32), wi_n_bad=np.zeros(W * nI, np.int32),
        slot_read=np.arange(W * R, dtype=np.int32), read_seq_off=np.arange(W * R + 1, dtype=np.int64) * read_len,
        read_seq=np.concatenate(read_seq_chunks) if W else np.zeros(0, np.uint8),
        read_qual=np.concatenate(read_qual_chunks) if W else np.zeros(0, np.uint8),
        read_pos=read_pos, read_end=(read_pos + read_len).astype(np.int32), read_mapq=read_mapq,
        read_qcfail=np.zeros(W * R, np.uint8))
    return b, VariantSet.from_lists(per_window_vars)


def _make_select_range(args):
    lo, hi, kw = args
    return make_select_batch(hi - lo, window_offset=lo, **kw)


def make_select_batch_parallel(n_windows, window_offset=0, n_procs=None, **kw):
    """make_select_batch() over a process pool (per-window RNG streams, so the bytes are identical)."""
    import multiprocessing as mp
    import os
    from .batch import VariantSet, concat_batches
    n_procs = n_procs or min(os.cpu_count() or 1, 32)
    if n_procs <= 1 or n_windows <= 512:
        return make_select_batch(n_windows, window_offset=window_offset, **kw)
    cuts = np.linspace(window_offset, window_offset + n_windows, 2 * n_procs + 1).astype(int)
    jobs = [(int(a), int(b), kw) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    with mp.get_context("fork").Pool(n_procs) as pool:
        parts = pool.map(_make_select_range, jobs)
    batch = concat_batches([p[0] for p in parts])
    vs = [p[1] for p in parts]
    nv = np.cumsum([0] + [len(v.var_pos) for v in vs])
    na = np.cumsum([0] + [int(v.var_added_off[-1]) for v in vs])
    return batch, VariantSet(
        np.concatenate([vs[0].win_var_off] + [v.win_var_off[1:] + nv[i] for i, v in enumerate(vs) if i]).astype(np.int32),
        np.concatenate([v.var_pos for v in vs]), np.concatenate([v.var_n_removed for v in vs]),
        np.concatenate([v.var_n_support for v in vs]),
        np.concatenate([vs[0].var_added_off] + [v.var_added_off[1:] + na[i] for i, v in enumerate(vs) if i]).astype(np.int64),
        np.concatenate([v.var_added[:int(v.var_added_off[-1])] for v in vs] + [np.zeros(1, np.uint8)]))
