"""
Platypus-shaped objects above the C-ABI seam: what `callVariantsInWindow`
(reference: src/cython/variantcaller.pyx:74-141) and the region loop around it (:566-615) talk to, with the
arithmetic behind them running in the CUDA library.

    pop = Population(options, engine)
    for window in windows:                       # the reference's loop, unchanged
        pop.reset()
        pop.setup(variants, haplotypes, genotypes, nInd, verbosity, readBuffers)
        pop.call(100, 1)
    for res in pop.flush():                      # the one thing that changes above the seam: windows accumulate and
        ...  res.genotypeLikelihoods, res.frequencies, ...          # are scored in ONE plb_population_run_host call

Same method names, argument meaning and error behaviour as the reference's classes:

* `Population.setup / call / reset` (src/cython/cpopulation.pxd:46-56, cpopulation.pyx:166-309, 678-720).  The fields
  `outputCallToVCF` reads back (variantcaller.pyx:582) - `genotypeLikelihoods`, `goodnessOfFitValues`, `frequencies`,
  `haplotypeIndexes`, `EMLikelihoods`, `genotypeCalls`, `variantPosteriors`, `varsByPos`, `nReads`,
  `maxLogLikelihoods` - are attributes of the per-window `WindowResult`; with `batch_windows=1` every `call()` flushes
  at once and the population object itself carries them, exactly like the reference.
* `Haplotype.alignReads / alignSingleRead` (src/cython/chaplotype.pxd:44-45, chaplotype.pyx:306-384): the
  likelihood array of one individual terminated by the reference's sentinel 999.
* `BamReadBuffer.setWindowPointers` (src/cython/cwindow.pyx:655-689) over `reads.ReadBuffer`.
* `getAllHLAHaplotypesInRegion` (src/cython/variantFilter.pyx:655-736): the --HLATyping haplotype selection, two batched
  scoring calls + the reference's heap bookkeeping.

Nothing here computes a likelihood: every number comes out of `Engine` (libplatypus_b200.so); without the library or a
GPU construction fails loudly.
"""
from typing import List, Optional, Sequence

import numpy as np

from . import _abi
from .batch import Read, Window, WindowBatch
from .reads import ReadBuffer, window_slice

SENTINEL = 999.0      # end-of-array marker of Haplotype.likelihoodCache, chaplotype.pyx:375


class PlatypusError(Exception):
    """Stands where the reference raises StandardError (cpopulation.pyx:209-225, chaplotype.pyx:180-183)."""


class Variant:
    """The fields of the reference's Variant (src/cython/variant.pyx:109-145) this path reads: refName, refPos, removed,
    added, nSupportingReads and the prior (Variant.calculatePrior, variant.pyx:219-259; candidate generation supplies it -
    the tandem-repeat error model behind it is out of scope).  Equality / hash / order as variant.pyx:282-353."""
    __slots__ = ("refName", "refPos", "removed", "added", "nSupportingReads", "prior", "varSource")

    def __init__(self, refName, refPos, removed, added, nSupportingReads=1, prior=None, varSource=1):
        self.refName, self.refPos = refName, int(refPos)
        self.removed, self.added = bytes(removed), bytes(added)
        self.nSupportingReads, self.prior = nSupportingReads, prior
        self.varSource = varSource       # PLATYPUS_VAR 1, FILE_VAR 2, ASSEMBLER_VAR 4 (variant.pyx:43-45)

    def _key(self):
        return (self.refName, self.refPos, self.removed, self.added)

    @property
    def varType(self):   # SNP 0, MNP 1, INS 2, DEL 3, REP 4 (variant.pyx:49-53, 136-144)
        na, nr = len(self.added), len(self.removed)
        if na == nr:
            return 0 if na == 1 else 1
        return 2 if nr == 0 else 3 if na == 0 else 4

    def _order(self):    # the reference sorts by (refName, refPos, varType, nRemoved), variant.pyx:304-315
        return (self.refName, self.refPos, self.varType, len(self.removed))

    def __eq__(self, other):
        return isinstance(other, Variant) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __lt__(self, other):
        return self._order() < other._order()

    def __repr__(self):
        return "%s:%d %s>%s" % (self.refName, self.refPos, self.removed.decode() or "-", self.added.decode() or "-")

    def calculatePrior(self):
        return 0.5 if self.prior is None else float(self.prior)


class Options:
    """The option fields the hot path reads (defaults of src/python/runner.py:519-597)."""

    def __init__(self, **kw):
        self.maxHaplotypes, self.maxVariants, self.maxReadLength, self.minPosterior = 50, 8, 150, 5
        self.useEMLikelihoods, self.calculateFlankScore, self.HLATyping, self.nInd, self.verbosity = 0, 0, 0, 1, 2
        self.rlen = 150
        self.originalMaxHaplotypes, self.coverageSamplingLevel = 50, 30
        for k, v in kw.items():
            setattr(self, k, v)

    def plb_options(self, maxIters=100):
        return _abi.PlbOptions.default(use_mapq_cap=int(bool(self.HLATyping)), calc_flank_score=int(bool(self.calculateFlankScore)),
                                       use_em_likelihoods=int(self.useEMLikelihoods), max_em_iters=int(maxIters))


def _as_read(r) -> Read:
    if isinstance(r, Read):
        return r
    if hasattr(r, "to_read"):
        return r.to_read()          # reads.AlignedRead
    seq, qual, pos, end, mapq, flag = r[:6]
    return Read(bytes(seq), bytes(qual), int(pos), int(end), int(mapq), bool(flag & 512) if isinstance(flag, int) else bool(flag))


class BamReadBuffer:
    """bamReadBuffer (src/cython/cwindow.pyx:485-768) as the window model sees it: after setWindowPointers the
    attributes `reads`, `badReads` and `brokenMates` hold the window's slices (cwindow.pyx:655-689, 208-236)."""

    def __init__(self, sample: str, buffer: Optional[ReadBuffer] = None, broken: Optional[list] = None):
        self.sample = sample
        self.buffer = buffer or ReadBuffer()
        self._broken = broken or []
        self.reads, self.badReads, self.brokenMates = [], [], []

    def setWindowPointers(self, start: int, end: int, *_ignored):
        self.reads = window_slice(self.buffer.reads, start, end)
        self.badReads = window_slice(self.buffer.bad_reads, start, end)
        self.brokenMates = window_slice(self._broken, start, end)


class WindowReads:
    """A window's three read lists of one individual, already sliced (lists of batch.Read / reads.AlignedRead / tuples
    (seq, qual, pos, end, mapq, bitFlag))."""

    def __init__(self, reads=(), badReads=(), brokenMates=(), sample=""):
        self.sample = sample
        self.reads, self.badReads, self.brokenMates = list(reads), list(badReads), list(brokenMates)


class Haplotype:
    """Haplotype (src/cython/chaplotype.pyx:127-191): window [startPos, endPos), a tuple of variants, the reference
    sequence of the window plus endBufferSize = min(2 * maxReadLength, 500) bases either side (chaplotype.pyx:142), and
    the mutated sequence built from it (chaplotype.pyx:397-449; on the GPU, plb_build_haplotypes_host).  `refFile` is
    anything with getSequence(refName, begin, end) (half-open, fastafile.pyx:173-207)."""

    def __init__(self, refName, startPos, endPos, variants, refFile, maxReadLength, options=None, engine=None,
                 haplotypeSequence: Optional[bytes] = None):
        self.refName, self.startPos, self.endPos = refName, int(startPos), int(endPos)
        self.variants = tuple(variants)
        self.options = options or Options()
        self.maxReadLength = int(maxReadLength)
        self.endBufferSize = min(2 * self.maxReadLength, 500)
        self.engine = engine
        lo = max(0, self.startPos - self.endBufferSize)
        self.referenceSequence = bytes(refFile.getSequence(refName, lo, self.endPos + self.endBufferSize)) if refFile is not None else None
        self.hapStart = lo
        self._seq = haplotypeSequence
        self.lastIndividualIndex = -1
        self.likelihoodCache = None
        if self._seq is not None and len(self._seq) > _abi.PLB_MAX_HAP_LEN:
            raise PlatypusError("Haplotype with vars %s has len %s. Start is %s. End is %s. maxReadLen = %s"
                                % (self.variants, len(self._seq), self.startPos, self.endPos, self.maxReadLength))

    # ---- sequence ----------------------------------------------------------------------------------------
    @staticmethod
    def build_sequences(haps: Sequence["Haplotype"], engine):
        """Builds the mutated sequences of many haplotypes in one plb_build_haplotypes_host call (haplotypes of one
        window share startPos / endPos / referenceSequence)."""
        from .batch import VariantSet
        todo = [h for h in haps if h._seq is None]
        if not todo:
            return
        wins, key_ix = [], {}
        for h in todo:
            k = (h.refName, h.startPos, h.endPos, h.hapStart)
            if k not in key_ix:
                key_ix[k] = len(wins)
                wins.append((h, []))
            vs = wins[key_ix[k]][1]
            for v in h.variants:
                if v not in vs:
                    vs.append(v)
        ref_wins, per_window = [], []
        for h, vs in wins:
            vs.sort()
            ref_wins.append(Window(h.startPos, h.endPos, h.hapStart, [h.referenceSequence], [([], [], [])]))
            per_window.append([(v.refPos, len(v.removed), v.added, v.nSupportingReads or 1) for v in vs])
        rb = WindowBatch.from_windows(ref_wins, 1)
        vset = VariantSet.from_lists(per_window)
        hap_win = [key_ix[(h.refName, h.startPos, h.endPos, h.hapStart)] for h in todo]
        masks = [sum(1 << wins[w][1].index(v) for v in h.variants) for h, w in zip(todo, hap_win)]
        for h, s in zip(todo, engine.build_haplotypes(rb, vset, hap_win, masks)):
            if len(s) > _abi.PLB_MAX_HAP_LEN:
                raise PlatypusError("Haplotype with vars %s has len %s" % (h.variants, len(s)))
            h._seq = s

    @property
    def haplotypeSequence(self) -> bytes:
        if self._seq is None:
            if not self.variants:
                self._seq = self.referenceSequence
            else:
                Haplotype.build_sequences([self], self._engine())
        return self._seq

    def _engine(self):
        if self.engine is None:
            raise PlatypusError("this Haplotype has no engine (pass engine= or score it through a Population)")
        return self.engine

    def __eq__(self, other):    # chaplotype.pyx:212-235: haplotypes compare by sequence
        return isinstance(other, Haplotype) and self.haplotypeSequence == other.haplotypeSequence

    def __hash__(self):
        return hash(self.haplotypeSequence)

    def _order(self):           # chaplotype.pyx:225-264: contig, window start, then the mutated sequence
        return (self.refName, self.startPos, self.haplotypeSequence)

    def __ne__(self, other):
        return not self.__eq__(other)

    def __lt__(self, other):
        return self._order() < other._order()

    def __gt__(self, other):
        return self._order() > other._order()

    def __le__(self, other):
        return not self._order() > other._order()

    def __ge__(self, other):
        return not self._order() < other._order()

    # ---- scoring (seam S2) -------------------------------------------------------------------------------
    def alignReads(self, individualIndex, reads, badReads, brokenReads, useMapQualCap=0):
        """Haplotype.alignReads (chaplotype.pyx:306-377): log-likelihood of every read of one individual - good, then
        bad, then broken mates - as one array terminated by the sentinel 999; cached per individualIndex like the
        reference's likelihoodCache (the returned array is borrowed: the next call with another index replaces it)."""
        if individualIndex == self.lastIndividualIndex and self.likelihoodCache is not None:
            return self.likelihoodCache
        g, b, k = [_as_read(r) for r in reads], [_as_read(r) for r in badReads], [_as_read(r) for r in brokenReads]
        w = Window(self.startPos, self.endPos, self.hapStart, [self.haplotypeSequence], [(g, b, k)])
        opt = self.options.plb_options()
        opt.use_mapq_cap = int(bool(useMapQualCap))
        ll, _ = self._engine().window_loglik(WindowBatch.from_windows([w], 1), opt=opt, want_score=False)
        self.likelihoodCache = np.concatenate([np.asarray(ll, np.float64), [SENTINEL]])
        self.lastIndividualIndex = individualIndex
        return self.likelihoodCache

    def alignSingleRead(self, theRead, useMapQualCap=0) -> float:
        """Haplotype.alignSingleRead (chaplotype.pyx:379-384): a bare alignReadToHaplotype - no QC-fail / overlap
        rule - which is how alignReads treats broken mates."""
        saved = (self.lastIndividualIndex, self.likelihoodCache)
        self.lastIndividualIndex, self.likelihoodCache = -1, None
        try:
            return float(self.alignReads(-2, [], [], [theRead], useMapQualCap)[0])
        finally:
            self.lastIndividualIndex, self.likelihoodCache = saved


PLATYPUS_VAR, FILE_VAR, ASSEMBLER_VAR = 1, 2, 4      # variant.pyx:43-45


def getAllHLAHaplotypesInRegion(chrom, windowStart, windowEnd, refFile, options, variants, refHaplotype, readBuffers,
                                engine=None):
    """getAllHLAHaplotypesInRegion (src/cython/variantFilter.pyx:655-736), the --HLATyping haplotype selection, with the
    reference's argument list (+ the engine): one haplotype per FILE_VAR variant; up to 150 of them are returned as they
    are.  Beyond that a heap keeps the originalMaxHaplotypes - 1 best (score, haplotype) tuples - first scored by
    computeBestScoreForHaplotype, the best 75 of the heap are output, then every haplotype is scored once more as a genotype
    with the best one (computeBestScoreForGenotype) and pushed onto the SAME heap, and the best 75 are output again (so the
    list may name a haplotype twice, as the reference's does).  The scores come from the GPU in two calls for the whole
    window (plb_best_score_haplotypes_host, plb_best_score_genotypes_host) instead of one alignSingleRead per read and
    haplotype; the heap bookkeeping is the reference's own, on the same tuple order (ties by Haplotype comparison)."""
    from heapq import heappush, heappushpop
    engine = engine or refHaplotype.engine
    maxReadLength = options.rlen
    allHaps = [Haplotype(chrom, windowStart, windowEnd, (v,), refFile, maxReadLength, options, engine)
               for v in variants if v.varSource == FILE_VAR]
    nHaps = len(allHaps)
    maxHaplotypes = 150                                # variantFilter.pyx:699
    if nHaps <= maxHaplotypes:
        return allHaps
    if engine is None:
        raise PlatypusError("getAllHLAHaplotypesInRegion needs an engine (there is no CPU path)")
    originalMaxHaplotypes = options.originalMaxHaplotypes - 1
    for k in range(0, nHaps, 64):                      # a variant mask holds 64 variants per build call
        Haplotype.build_sequences(allHaps[k:k + 64], engine)
    per_ind = [([_as_read(r) for r in rb.reads], [], []) for rb in readBuffers]
    win = Window(int(windowStart), int(windowEnd), allHaps[0].hapStart, [h.haplotypeSequence for h in allHaps], per_ind)
    batch = WindowBatch.from_windows([win], len(per_ind), dedupe_reads=False)
    opt = options.plb_options()
    hapsByBestScore, outputHaps = [], []

    def push(item):
        if len(hapsByBestScore) < originalMaxHaplotypes:
            heappush(hapsByBestScore, item)
        else:
            heappushpop(hapsByBestScore, item)
    scores = engine.best_score_haplotypes(batch, opt=opt)
    for k in range(nHaps):
        push((float(scores[k]), allHaps[k]))
    for index, (score, thisHap) in enumerate(sorted(hapsByBestScore, reverse=True)):
        if index < maxHaplotypes / 2:
            outputHaps.append(thisHap)
        else:
            break
    bestHap = sorted(hapsByBestScore, reverse=True)[0][1]
    best = next(k for k, h in enumerate(allHaps) if h is bestHap)
    scores = engine.best_score_genotypes(batch, [best] * nHaps, list(range(nHaps)), options.coverageSamplingLevel, opt=opt)
    for k in range(nHaps):
        push((float(scores[k]), allHaps[k]))
    for index, (score, thisHap) in enumerate(sorted(hapsByBestScore, reverse=True)):
        if index < maxHaplotypes / 2:
            outputHaps.append(thisHap)
        else:
            break
    return outputHaps


def generateAllGenotypesFromHaplotypeList(haplotypes):
    """Genotype order of cgenotype.pyx:193-218: (i, j) for i in 0..H, j in i..H.  Returns (hap_i, hap_j) pairs."""
    n = len(haplotypes)
    return [(haplotypes[i], haplotypes[j]) for i in range(n) for j in range(i, n)]


class WindowResult:
    """The fields of Population that variantcaller.pyx:582 hands to outputCallToVCF, for one window."""

    def __init__(self, variants, haplotypes, genotypes, nInd, readBuffers):
        self.variants, self.haplotypes, self.genotypes = list(variants), list(haplotypes), list(genotypes)
        self.nIndividuals, self.readBuffers = nInd, readBuffers
        self.nHaplotypes, self.nGenotypes = len(self.haplotypes), len(self.genotypes)
        self.nVariants = len(self.variants)
        self.genotypeLikelihoods = self.goodnessOfFitValues = self.EMLikelihoods = None
        self.frequencies = self.haplotypeIndexes = self.nReads = self.maxLogLikelihoods = None
        self.genotypeCalls, self.variantPosteriors, self.varsByPos = [], {}, {}
        self.vcfInfo, self.vcfFilter = {}, {}
        self.emIterations = 0


class Population:
    """Population (src/cython/cpopulation.pyx:82-720) with the window model on the GPU.  `batch_windows` windows are
    accumulated between flushes (1 = the reference's behaviour: call() returns with the results on `self`)."""

    _FIELDS = ("variants", "haplotypes", "genotypes", "nIndividuals", "nHaplotypes", "nGenotypes", "nVariants",
               "genotypeLikelihoods", "goodnessOfFitValues", "EMLikelihoods", "frequencies", "haplotypeIndexes", "nReads",
               "maxLogLikelihoods", "genotypeCalls", "variantPosteriors", "varsByPos", "vcfInfo", "vcfFilter", "readBuffers")

    def __init__(self, options=None, engine=None, batch_windows: int = 4096):
        from .engine import Engine
        self.options = options or Options()
        self.engine = engine or Engine(0)
        self.maxHaplotypes = int(self.options.maxHaplotypes)
        self.maxGenotypes = self.maxHaplotypes * (self.maxHaplotypes + 1) // 2     # variantcaller.pyx:916-924
        self.useEMLikelihoods = int(self.options.useEMLikelihoods)
        self.batch_windows = max(1, int(batch_windows))
        self._pending: List[WindowResult] = []
        self._current: Optional[WindowResult] = None
        self._called: List[tuple] = []
        self.reset()

    def reset(self):
        """Population.reset (cpopulation.pyx:166-195): forget the current window (queued windows are kept)."""
        self._current = None
        for f in self._FIELDS:
            setattr(self, f, None)

    def setup(self, variants, haplotypes, genotypes, nInd, verbosity, readBuffers):
        """Population.setup (cpopulation.pyx:197-309): same limits, same exception text."""
        nH, nG = len(haplotypes), len(genotypes)
        if nInd != len(readBuffers):
            raise PlatypusError("Number of individuals (%s) does not match number of read buffers (%s)" % (nInd, len(readBuffers)))
        if nH > self.maxHaplotypes:
            raise PlatypusError("Too many haplotypes. %s > %s" % (nH, self.maxHaplotypes))
        if nG > self.maxGenotypes:
            raise PlatypusError("Too many genotypes. %s > %s" % (nG, self.maxGenotypes))
        if nG != nH * (nH + 1) // 2:
            raise PlatypusError("genotypes must be generateAllGenotypesFromHaplotypeList(haplotypes): %s != %s" % (nG, nH * (nH + 1) // 2))
        res = WindowResult(variants, haplotypes, genotypes, nInd, readBuffers)
        # snapshot the window's read slices now: the buffers' window pointers move on with the caller's loop
        res._reads = [([_as_read(r) for r in rb.reads], [_as_read(r) for r in rb.badReads], [_as_read(r) for r in rb.brokenMates])
                      for rb in readBuffers]
        self._current = res

    def call(self, maxIters=100, computeVCFFields=0):
        """Population.call (cpopulation.pyx:678-720).  Queues the window; with batch_windows == 1 (or once that many
        windows are queued) the batch is flushed.  computeVCFFields is accepted for signature compatibility - INFO /
        FILTER text is the VCF writer's (out of scope)."""
        if self._current is None:
            raise PlatypusError("call() before setup()")
        res = self._current
        res._max_iters = int(maxIters)
        self._pending.append(res)
        if len(self._pending) >= self.batch_windows:
            out = self.flush()
            if self.batch_windows == 1:
                for f in self._FIELDS:
                    setattr(self, f, getattr(out[-1], f))
            else:
                self._called.extend(out)

    def flush(self) -> List[WindowResult]:
        """Scores every queued window in one plb_population_run_host call per (nInd, maxIters) group and fills the
        WindowResult objects; returns them in the order of their call()s (including any flushed early)."""
        done, self._called = self._called, []
        pend, self._pending = self._pending, []
        groups = {}
        for r in pend:
            groups.setdefault((r.nIndividuals, r._max_iters), []).append(r)
        for (nInd, iters), rs in groups.items():
            Haplotype.build_sequences([h for r in rs for h in r.haplotypes], self.engine)
            wins = []
            for r in rs:
                h0 = r.haplotypes[0]
                if any((h.startPos, h.endPos, h.hapStart) != (h0.startPos, h0.endPos, h0.hapStart) for h in r.haplotypes):
                    raise PlatypusError("the haplotypes of a window must share its interval and flank")
                vs = r.variants
                masks = [sum(1 << vs.index(v) for v in h.variants if v in vs) for h in r.haplotypes]
                wins.append(Window(h0.startPos, h0.endPos, h0.hapStart, [h.haplotypeSequence for h in r.haplotypes], r._reads,
                                   hap_var_mask=masks, var_prior=[v.calculatePrior() for v in vs] or None))
            batch = WindowBatch.from_windows(wins, nInd)
            out = self.engine.population_run(batch, opt=self.options.plb_options(iters))
            for w, r in enumerate(rs):
                self._fill(r, out, w)
        order = {id(r): i for i, r in enumerate(pend)}
        return done + sorted(pend, key=lambda r: order[id(r)])

    def _fill(self, r: WindowResult, out, w):
        H, G, nI = r.nHaplotypes, r.nGenotypes, r.nIndividuals
        r.genotypeLikelihoods = out["gl"][w, :, :G].copy()                  # [nInd][G], cpopulation.pyx:304-309
        r.goodnessOfFitValues = out["gof"][w, :G, :].copy()                 # [G][nInd]
        r.EMLikelihoods = out["em_post"][w, :, :G].copy()
        r.frequencies = out["freq"][w, :H].copy()
        r.maxLogLikelihoods = out["gl_log_max"][w].copy()
        r.haplotypeIndexes = np.array([(i, j) for i in range(H) for j in range(i, H)], np.int32).reshape(G, 2)
        r.nReads = np.array([len(g) for g, _, _ in r._reads], np.int32)     # good reads only, cpopulation.pyx:286-287
        r.emIterations = int(out["em_iters"][w])
        r.genotypeCalls = [None if c < 0 else r.genotypes[int(c)] for c in out["call"][w]]    # cpopulation.pyx:623-676
        r.variantPosteriors, r.varsByPos = {}, {}
        done = set()
        for hap in r.haplotypes:                                             # cpopulation.pyx:596-621, same visiting order
            for v in hap.variants:
                if v in done or v not in r.variants:
                    continue
                done.add(v)
                post = float(out["var_phred"][w, r.variants.index(v)])
                if post >= self.options.minPosterior:
                    r.variantPosteriors[v] = post
                    r.varsByPos.setdefault(v.refPos, []).append(v)
