"""
CPU restatement of the reference's haplotype construction and haplotype selection loop.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg, never by
platypus_b200/.  Small-case oracle in Python on purpose: the reference's loop is Python too, so `heapq`,
`sorted`, tuple comparison and `itertools.combinations` here ARE the algorithms the reference runs, and math.exp /
math.log are the libm calls its C code makes.

Follows
    Variant.__init__ / __richcmp__            src/cython/variant.pyx:109-145, 282-353
    isHaplotypeValid                          src/cython/platypusutils.pyx:735-802
    Haplotype.__init__ / getMutatedSequence   src/cython/chaplotype.pyx:127-191, 397-449
    computeBestScoreForHaplotype              src/cython/variantFilter.pyx:212-234
    computeBestScoreForGenotype               src/cython/variantFilter.pyx:237-283
    getFilteredHaplotypes                     src/cython/variantFilter.pyx:377-506
Per-read log-likelihoods come from the C oracle (plo_window_loglik; the sampled reads are passed as broken mates, the
list Haplotype.alignReads scores with a bare alignReadToHaplotype - exactly alignSingleRead, chaplotype.pyx:368-372,
379-384).

Pinned against the reference's own functions (oracle/_ref/n1_ref, built by oracle/build.py): tests/test_oracle.py
test_n1_*; golden vectors tests/golden/n1_ref.npz.
"""
import heapq
import math
from itertools import combinations

from platypus_b200.batch import Read, Window, WindowBatch

SNP, MNP, INS, DEL, REP = 0, 1, 2, 3, 4


class Var:
    """The fields and the ordering of the reference's Variant (variant.pyx:109-145, 282-353)."""
    __slots__ = ("idx", "pos", "removed", "added", "n_support", "n_added", "n_removed", "min_pos", "max_pos", "vtype")

    def __init__(self, idx, pos, removed, added, n_support):
        pos = max(0, pos)
        self.idx, self.pos, self.removed, self.added, self.n_support = idx, pos, removed, added, n_support
        self.n_added, self.n_removed = len(added), len(removed)
        self.min_pos = pos
        self.max_pos = max(pos, pos + self.n_removed - 1)
        if self.n_removed == self.n_added:
            self.vtype = SNP if self.n_added == 1 else MNP
        elif self.n_removed == 0:
            self.vtype = INS
        elif self.n_added == 0:
            self.vtype = DEL
        else:
            self.vtype = REP

    def _key(self):
        return (self.pos, self.vtype, self.n_removed)

    def __lt__(self, o):
        return self._key() < o._key()

    def __gt__(self, o):
        return self._key() > o._key()

    def __le__(self, o):
        return not self._key() > o._key()

    def __ge__(self, o):
        return not self._key() < o._key()

    def __eq__(self, o):
        return self.pos == o.pos and self.added == o.added and self.removed == o.removed

    def __ne__(self, o):
        return not self == o

    def __hash__(self):
        return hash((self.pos, self.removed, self.added))


def is_haplotype_valid(vs):
    """platypusutils.pyx:735-802."""
    n = len(vs)
    if n <= 1:
        return True
    for i in range(n - 1):
        a, b = vs[i], vs[i + 1]
        if a.min_pos > b.min_pos:
            raise ValueError("Variants out of order in haplotype!")
        if a.max_pos > b.min_pos:
            return False
        if a.max_pos == b.min_pos:
            if a.n_added == a.n_removed and b.n_added != b.n_removed:
                continue
            return False
    return True


def build_haplotype(ref_seq, win_start, win_end, hap_start, vs):
    """Haplotype.haplotypeSequence (chaplotype.pyx:163-172, 397-449).  ref_seq = Haplotype.referenceSequence of the
    window = refFile.getSequence(startPos - endBufferSize, endPos + endBufferSize), whose first base sits at
    max(0, hap_start) (fastafile.pyx:173-207 clamps the interval to the contig)."""
    if not vs:
        return ref_seq
    ref_start = win_start - min(win_start - hap_start, win_start)

    def seq(a, b):          # refFile.getSequence inside the window's segment
        return ref_seq[a - ref_start:b - ref_start]
    cur = win_start
    bits = []
    first = vs[0]
    if first.pos != cur:
        bits.append(seq(cur, first.pos))
        cur = first.pos
    for v in vs:
        if v.pos > cur:
            bits.append(seq(cur, v.pos))
            cur = v.pos
        if v.n_added == v.n_removed:
            bits.append(v.added)
            cur += v.n_removed
        else:
            if v.n_added == 0 or v.n_removed == 0:
                if v.pos == cur:
                    bits.append(seq(v.pos, v.pos + 1))
                    cur += 1
            cur += v.n_removed
            bits.append(v.added)
    if cur < win_end:
        bits.append(seq(cur, win_end))
    return seq(ref_start, win_start) + b"".join(bits) + ref_seq[win_end - ref_start:]


class SelectWindow:
    """Inputs of getFilteredHaplotypes for one window."""

    def __init__(self, ref_seq, win_start, win_end, hap_start, variants, good_reads):
        """variants: [(refPos, removed, added, nSupportingReads)] in the window's order; good_reads: per individual
        the reads.windowStart..windowEnd list of platypus_b200.batch.Read."""
        self.ref_seq, self.win_start, self.win_end, self.hap_start = ref_seq, win_start, win_end, hap_start
        self.vars = [Var(i, p, rem, add, n) for i, (p, rem, add, n) in enumerate(variants)]
        self.good = good_reads


def _sampled(w, target_coverage):
    """The reads computeBestScoreForGenotype visits (variantFilter.pyx:253-277): every sampleRate-th good read."""
    out = []
    size = w.win_end - w.win_start
    for reads in w.good:
        if not reads:
            out.append([])
            continue
        mean_cov = reads[0].rlen * len(reads) // size
        rate = max(1, mean_cov // target_coverage)
        out.append(list(reads[::rate]))
    return out


def _loglik(w, sampled, hap_seqs, opt):
    """alignSingleRead(read, False) of every sampled read against every sequence: ll[h][individual][k]."""
    from . import oracle as O
    win = Window(w.win_start, w.win_end, w.hap_start, list(hap_seqs), [([], [], list(s)) for s in sampled])
    b = WindowBatch.from_windows([win], len(sampled), dedupe_reads=False)
    ll, _, _ = O.window_loglik(b, opt, want_score=False)
    off = b.ll_offsets()
    H = len(hap_seqs)
    out = [[None] * len(sampled) for _ in range(H)]
    for i, s in enumerate(sampled):
        T = len(s)
        for h in range(H):
            out[h][i] = ll[off[i] + h * T: off[i] + (h + 1) * T]
    return out


def best_scores(w, var_sets, target_coverage=30, opt=None):
    """computeBestScoreForGenotype(readBuffers, DiploidGenotype(ref, Haplotype(set)), windowSize, targetCoverage) for
    every variant set (tuples of Var)."""
    sampled = _sampled(w, target_coverage)
    seqs = [w.ref_seq] + [build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start, vs) for vs in var_sets]
    ll = _loglik(w, sampled, seqs, opt)
    scores = []
    for k in range(len(var_sets)):
        best = -1e20
        for i, s in enumerate(sampled):
            if not w.good[i]:
                continue
            tot = 0.0
            for t in range(len(s)):
                tot += math.log(0.5 * (math.exp(ll[0][i][t]) + math.exp(ll[k + 1][i][t])))
            best = max(best, tot)
        scores.append(best)
    return scores


def best_score_haplotypes(w, var_sets, opt=None):
    """computeBestScoreForHaplotype(readBuffers, Haplotype(set)) (variantFilter.pyx:212-234) for every variant set: the
    sum of alignSingleRead over ALL good reads of an individual, best individual (one without reads sums to 0.0)."""
    reads = [list(r) for r in w.good]
    seqs = [build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start, vs) for vs in var_sets]
    ll = _loglik(w, reads, seqs, opt) if seqs else []
    out = []
    for k in range(len(var_sets)):
        best = -1e20
        for i, r in enumerate(reads):
            tot = 0.0
            for t in range(len(r)):
                tot += ll[k][i][t]
            best = max(best, tot)
        out.append(best)
    return out


def best_score_genotype_pairs(w, sets1, sets2, target_coverage=30, opt=None):
    """computeBestScoreForGenotype(readBuffers, DiploidGenotype(Haplotype(s1), Haplotype(s2)), windowSize, targetCoverage)
    (variantFilter.pyx:237-283) for every pair of variant sets."""
    sampled = _sampled(w, target_coverage)
    uniq = {}
    for vs in list(sets1) + list(sets2):
        uniq.setdefault(tuple(v.idx for v in vs), vs)
    keys = list(uniq)
    seqs = [build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start, uniq[k]) for k in keys]
    ll = _loglik(w, sampled, seqs, opt) if seqs else []
    at = {k: i for i, k in enumerate(keys)}
    scores = []
    for a, b in zip(sets1, sets2):
        ia, ib = at[tuple(v.idx for v in a)], at[tuple(v.idx for v in b)]
        best = -1e20
        for i, smp in enumerate(sampled):
            if not w.good[i]:
                continue
            tot = 0.0
            for t in range(len(smp)):
                tot += math.log(0.5 * (math.exp(ll[ia][i][t]) + math.exp(ll[ib][i][t])))
            best = max(best, tot)
        scores.append(best)
    return scores


class _HapKey:
    """Order and equality of the reference's Haplotype objects inside its (score, haplotype) tuples
    (chaplotype.pyx:218-287): same contig and window start here, so the sequence decides."""
    __slots__ = ("seq", "idx")

    def __init__(self, seq, idx):
        self.seq, self.idx = seq, idx

    def __lt__(self, o):
        return self.seq < o.seq

    def __gt__(self, o):
        return self.seq > o.seq

    def __le__(self, o):
        return not self.seq > o.seq

    def __ge__(self, o):
        return not self.seq < o.seq

    def __eq__(self, o):
        return self.seq == o.seq

    def __ne__(self, o):
        return self.seq != o.seq

    __hash__ = None


def hla_haplotypes(w, sources, original_max_haplotypes=50, target_coverage=30, opt=None, scores=None):
    """getAllHLAHaplotypesInRegion (variantFilter.pyx:655-736): one haplotype per FILE_VAR variant (sources[i] == 2);
    up to 150 of them are returned as they are, otherwise a heap keeps the originalMaxHaplotypes - 1 best by
    computeBestScoreForHaplotype, the best 75 of the heap are output, every haplotype is scored once more as a genotype
    with the best one, pushed onto the SAME heap, and the best 75 of the heap are output again.
    Returns the variant index of every returned haplotype, in order.  scores = (hap_scores, gt_score_fn) replaces the
    oracle's own scoring (gt_score_fn(best position in the FILE_VAR list) -> genotype scores)."""
    from heapq import heappush, heappushpop
    file_vars = [i for i, s_ in enumerate(sources) if s_ == 2]
    n = len(file_vars)
    max_haps = 150                                    # variantFilter.pyx:699 overrides options.maxHaplotypes
    if n <= max_haps:
        return list(file_vars)
    cap = original_max_haplotypes - 1
    sets = [(w.vars[i],) for i in file_vars]
    keys = [_HapKey(build_haplotype(w.ref_seq, w.win_start, w.win_end, w.hap_start, vs), i) for vs, i in zip(sets, file_vars)]
    hs = scores[0] if scores else best_score_haplotypes(w, sets, opt)
    heap, out = [], []

    def push(item):
        if len(heap) < cap:
            heappush(heap, item)
        else:
            heappushpop(heap, item)
    for k in range(n):
        push((hs[k], keys[k]))
    for index, (sc, key) in enumerate(sorted(heap, reverse=True)):
        if index < max_haps / 2:
            out.append(key.idx)
        else:
            break
    best = sorted(heap, reverse=True)[0][1]
    pos = file_vars.index(best.idx)
    gs = scores[1](pos) if scores else best_score_genotype_pairs(w, [sets[pos]] * n, sets, target_coverage, opt)
    for k in range(n):
        push((gs[k], keys[k]))
    for index, (sc, key) in enumerate(sorted(heap, reverse=True)):
        if index < max_haps / 2:
            out.append(key.idx)
        else:
            break
    return out


def select_haplotypes(w, max_haplotypes=50, original_max_haplotypes=50, max_variants=8, filter_by_coverage=1,
                      target_coverage=30, opt=None, trace=None, score_fn=None):
    """getFilteredHaplotypes (variantFilter.pyx:377-506).  Returns the variant-index tuple of every haplotype it
    returns, in order, and the score it was kept with (None in the enumerate-everything branch)."""
    orig_cap = original_max_haplotypes - 1
    cap = max_haplotypes - 1
    n = len(w.vars)
    if n <= math.log2(cap) or (filter_by_coverage and max_variants <= math.log2(cap)):
        out = []
        for k in range(1, n + 1):
            for vs in combinations(w.vars, k):
                if is_haplotype_valid(vs):
                    out.append((tuple(v.idx for v in vs), None))
        return out
    by_cov = sorted(w.vars, key=lambda v: v.n_support, reverse=True)
    heap = []
    for tv in by_cov:
        old = sorted(heap)
        trials = [(tv,)]
        for _, vs2 in old:
            both = tuple(sorted((tv,) + vs2))
            if is_haplotype_valid(both):
                trials.append(both)
        # scoring a trial does not depend on the heap, so the round's trials are scored together; the heap
        # operations below then run in the reference's order
        # score_fn (tests of the bookkeeping alone): a stand-in for computeBestScoreForGenotype, called per variant tuple
        scores = [score_fn(vs) for vs in trials] if score_fn else best_scores(w, trials, target_coverage, opt)
        if trace is not None:
            trace.append([(tuple(v.idx for v in vs), s) for vs, s in zip(trials, scores)])
        for vs, s in zip(trials, scores):
            if len(heap) < orig_cap:
                heapq.heappush(heap, (s, vs))
            else:
                heapq.heappushpop(heap, (s, vs))
    out = []
    for i, (s, vs) in enumerate(sorted(heap, reverse=True)):
        if i < cap:
            out.append((tuple(v.idx for v in vs), s))
        else:
            break
    return out


def window_from_batch(batch, vset, w):
    """SelectWindow of window w of a (WindowBatch with one reference haplotype per window, VariantSet) pair - the
    inputs of plb_select_haplotypes_host."""
    nI = batch.n_individuals
    vs = []
    for i in range(int(vset.win_var_off[w]), int(vset.win_var_off[w + 1])):
        add = vset.var_added[int(vset.var_added_off[i]):int(vset.var_added_off[i + 1])].tobytes()
        vs.append((int(vset.var_pos[i]), b"N" * int(vset.var_n_removed[i]), add, int(vset.var_n_support[i])))
    good = []
    for i in range(nI):
        wi = w * nI + i
        s0 = int(batch.wi_slot_off[wi])
        reads = []
        for s in range(s0, s0 + int(batch.wi_n_good[wi])):
            r = int(batch.slot_read[s])
            a, b = int(batch.read_seq_off[r]), int(batch.read_seq_off[r + 1])
            reads.append(Read(batch.read_seq[a:b].tobytes(), batch.read_qual[a:b].tobytes(), int(batch.read_pos[r]),
                              int(batch.read_end[r]), int(batch.read_mapq[r]), bool(batch.read_qcfail[r])))
        good.append(reads)
    h = int(batch.win_hap_off[w])
    ref = batch.hap_seq[int(batch.hap_seq_off[h]):int(batch.hap_seq_off[h + 1])].tobytes()
    return SelectWindow(ref, int(batch.win_start[w]), int(batch.win_end[w]), int(batch.hap_start[w]), vs, good)
