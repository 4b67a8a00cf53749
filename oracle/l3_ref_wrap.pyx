# cython: language_level=3, boundscheck=False, wraparound=False
"""
Python entry points onto the REFERENCE's own src/cython/chaplotype.pyx and cgenotype.pyx (built in a
scratch directory by oracle/build.py with the non-algorithmic accommodations listed there), so that tests
can pin the oracle's restatement of
    Haplotype.alignReads / alignReadToHaplotype       (chaplotype.pyx:306-377, 594-676)
    DiploidGenotype.calculateDataLikelihood            (cgenotype.pyx:131-189)
against the reference itself.

TEST INFRASTRUCTURE ONLY.  This file contains no algorithm: it serves reference sequence from memory
(the reference reads it from a FASTA file), builds the objects and read-pointer arrays the reference
expects and forwards the calls.  Only built when /root/reference is present; the result lives in
oracle/_ref/.
"""
from libc.stdlib cimport malloc, calloc, free
from libc.string cimport memcpy
from cpython.exc cimport PyErr_Clear

cimport fastafile
cimport variant
cimport chaplotype
cimport cgenotype
from fastafile cimport FastaFile
import fastafile as _fastafile_module   # sequenceTuple is not in the .pxd
from variant cimport Variant
from chaplotype cimport Haplotype
from cgenotype cimport DiploidGenotype
from htslibWrapper cimport cAlignedRead
cimport cwindow
cimport cpopulation
from cwindow cimport bamReadBuffer
from cpopulation cimport Population
from cgenotype cimport generateAllGenotypesFromHaplotypeList


cdef class MemFasta(FastaFile):
    """A FastaFile whose one sequence lives in memory.  getSequence / getCharacter return what the
    reference's file-backed versions return for a FASTA holding the same bases (half-open interval, the
    same clamps: fastafile.pyx:173-207, 119-139)."""
    cdef bytes genome

    def __init__(self, bytes name, bytes genome):
        self.genome = genome.upper()
        self.refs = {name: _fastafile_module.sequenceTuple(name, len(genome), 0, len(genome), len(genome) + 1)}
        self.cache = None

    cdef bytes getSequence(self, bytes seqName, long long int beginPos, long long int endPos):
        cdef long long int seqLength = len(self.genome)
        beginPos = max(0, beginPos)
        endPos = min(seqLength - 1, endPos)
        if endPos < beginPos:
            raise IndexError("Cannot have beginPos = %s, endPos = %s" % (beginPos, endPos))
        return self.genome[beginPos:endPos]

    cdef bytes getCharacter(self, bytes seqName, long long int pos):
        if pos >= len(self.genome) or pos < 0:
            return b"-"
        return self.genome[pos:pos + 1]


class _Options(object):
    """The attributes of the reference's `options` object that Haplotype, bamReadBuffer and Population read
    (defaults of src/python/runner.py:519-597)."""
    def __init__(self, flank, hla=0, n_ind=1, max_haps=64, use_em=0):
        self.verbosity = 0
        self.calculateFlankScore = flank
        self.HLATyping = hla
        self.nInd = n_ind
        self.maxHaplotypes = max_haps
        self.maxGenotypes = max_haps * (max_haps + 1) // 2
        self.useEMLikelihoods = use_em
        self.minPosterior = -1            # keep every variant's posterior (the threshold is applied by the caller)
        self.rlen = 150
        self.maxReads = 5000000
        self.minBaseQual = 20
        self.minFlank = 10
        self.trimReadFlank = 0
        self.minMapQual = 20
        self.minGoodQualBases = 20
        self.trimOverlapping = 1
        self.trimAdapter = 1
        self.trimSoftClipped = 1
        self.filterDuplicates = 1
        self.filterReadsWithUnmappedMates = 1
        self.filterReadsWithDistantMates = 1
        self.filterReadPairsWithSmallInserts = 1


cdef cAlignedRead** _make_reads(list reads, list keep, int shift=0) except NULL:
    cdef int n = len(reads)
    cdef cAlignedRead** arr = <cAlignedRead**>calloc(n + 1, sizeof(cAlignedRead*))
    cdef cAlignedRead* r
    cdef bytes seq, qual
    cdef int i
    for i in range(n):
        seq, qual, pos, end, mapq, flag = reads[i]
        keep.append(seq)
        keep.append(qual)
        r = <cAlignedRead*>calloc(1, sizeof(cAlignedRead))
        r.seq = <char*>seq
        r.qual = <char*>qual
        r.rlen = len(seq)
        r.pos = pos - shift
        r.end = end - shift
        r.mapq = mapq
        r.bitFlag = flag
        r.hash = NULL
        arr[i] = r
    return arr


cdef void _free_reads(cAlignedRead** arr, int n):
    cdef int i
    for i in range(n):
        if arr[i].hash != NULL:
            free(arr[i].hash)
        free(arr[i])
    free(arr)


def window_likelihoods(bytes genome, int win_start, int win_end, list hap_variants, list good, list bad, list broken,
                       int max_read_len=150, int hla=0, int flank=0):
    """One window through the reference's Haplotype / DiploidGenotype classes.

    genome        the reference sequence (positions are indices into it)
    hap_variants  per haplotype: list of (refPos, removed, added), sorted by position ([] = reference haplotype)
    good/bad/broken  lists of (seq, qual (raw phred bytes), pos, end, mapq, bitFlag)
    Returns dict: hap_seq (Haplotype.cHaplotypeSequence), hap_start (startPos - endBufferSize),
    ll[h][t] (Haplotype.alignReads, reads in good|bad|broken order) and genotypes
    [(i, j, logLikelihood, gof, hap1Like, hap2Like)] for i <= j (DiploidGenotype.calculateDataLikelihood).
    """
    cdef bytes name = b"chr"
    cdef MemFasta fa = MemFasta(name, genome)
    opts = _Options(flank)
    cdef list keep = []
    cdef int ng = len(good), nb = len(bad), nk = len(broken)
    cdef cAlignedRead** g = _make_reads(good, keep)
    cdef cAlignedRead** b = _make_reads(bad, keep)
    cdef cAlignedRead** k = _make_reads(broken, keep)
    cdef list haps = []
    cdef Haplotype hap, hap2
    cdef Variant v
    cdef double* arr
    cdef double gof[1]
    cdef double logl
    cdef DiploidGenotype gt
    cdef int t, i, j, T = ng + nb + nk
    out = {"hap_seq": [], "ll": [], "genotypes": [], "hap_start": None}
    try:
        for vs in hap_variants:
            variants = tuple(Variant(name, p, rem, add, 1, 1) for (p, rem, add) in vs)
            hap = Haplotype(name, win_start, win_end, variants, fa, max_read_len, opts)
            haps.append(hap)
            out["hap_seq"].append(<bytes>hap.cHaplotypeSequence[:hap.hapLen])
            out["hap_start"] = hap.startPos - hap.endBufferSize
        for hap in haps:
            arr = hap.alignReads(0, g, g + ng, b, b + nb, k, k + nk, hla)
            out["ll"].append([arr[t] for t in range(T)])
            assert arr[T] == 999
        for i in range(len(haps)):
            for j in range(i, len(haps)):
                hap = haps[i]
                hap2 = haps[j]
                gt = DiploidGenotype(hap, hap2)
                gof[0] = 0.0
                logl = gt.calculateDataLikelihood(g, g + ng, b, b + nb, k, k + nk, 0, 1, gof, hla)
                out["genotypes"].append((i, j, logl, gof[0], gt.hap1Like, gt.hap2Like))
    finally:
        haps = []
        _free_reads(g, ng)
        _free_reads(b, nb)
        _free_reads(k, nk)
    return out


def population(bytes genome, int win_start, int win_end, list hap_variants, list per_ind_reads, int max_read_len=150,
               int hla=0, int flank=0, int use_em=0, int max_iters=100):
    """One window through the reference's Population class (src/cython/cpopulation.pyx): setup() then
    call(max_iters, 0).  per_ind_reads: per individual (good, bad, broken) read lists as in
    window_likelihoods.  Reads go straight into the buffers' arrays (no filtering: the lists ARE the window's
    reads) and the window pointers span them.
    Returns dict: hap_seq, hap_start, freq [H], gl [nInd][G] (rescaled), em [nInd][G], gl_log_max [nInd],
    gof [G][nInd], call [nInd] (genotype index, -1 = no reads), variants [(refPos, removed, added,
    phred posterior under the flat prior 0.5, Variant.calculatePrior or None, phred posterior under it or
    None, [haplotype indices holding it])]."""
    cdef bytes name = b"chr"
    cdef MemFasta fa = MemFasta(name, genome)
    cdef int n_ind = len(per_ind_reads)
    cdef int H = len(hap_variants)
    opts = _Options(flank, hla, n_ind, max(H, 2), use_em)
    cdef list keep = [], haps = [], buffers = [], all_arrays = []
    cdef Haplotype hap
    cdef bamReadBuffer buf
    cdef Population pop
    cdef cAlignedRead** arr
    cdef Variant v
    cdef int i, g, h, k, n, G = H * (H + 1) // 2
    cdef dict uniq = {}
    out = {"hap_seq": [], "hap_start": None}
    try:
        for vs in hap_variants:
            variants = []
            for (p, rem, add) in vs:      # one Variant object per distinct variant, shared between haplotypes
                key = (p, rem, add)
                if key not in uniq:
                    uniq[key] = Variant(name, p, rem, add, 1, 1)
                variants.append(uniq[key])
            hap = Haplotype(name, win_start, win_end, tuple(variants), fa, max_read_len, opts)
            haps.append(hap)
            out["hap_seq"].append(<bytes>hap.cHaplotypeSequence[:hap.hapLen])
            out["hap_start"] = hap.startPos - hap.endBufferSize
        for (good, bad, broken) in per_ind_reads:
            buf = bamReadBuffer(name, win_start, win_end, opts)
            buf.sample = b"s"
            for lst, ra in ((good, buf.reads), (bad, buf.badReads), (broken, buf.brokenMates)):
                n = len(lst)
                arr = _make_reads(lst, keep)
                all_arrays.append((<size_t>arr, n))
                for k in range(n):
                    (<cwindow.ReadArray>ra).append(arr[k])
                (<cwindow.ReadArray>ra).windowStart = (<cwindow.ReadArray>ra).array
                (<cwindow.ReadArray>ra).windowEnd = (<cwindow.ReadArray>ra).array + n
            buffers.append(buf)
        genotypes = generateAllGenotypesFromHaplotypeList(haps)
        pop = Population(opts)
        pop.setup(list(uniq.values()), haps, genotypes, n_ind, 0, buffers)
        try:
            pop.call(max_iters, 0)
        except Exception:
            pass
        # call() ends with computeVariantPosteriors, which asks every variant for its prior; for indels that is
        # the indel prior model of variant.pyx, which indexes Python-2 strings as char* and fails under Python 3
        # AFTER the EM and the genotype calls are complete.  The prior model is outside this path (priors enter as
        # numbers), so a pending error from it is dropped and the posteriors are taken below with explicit priors.
        PyErr_Clear()
        out["freq"] = [pop.frequencies[h] for h in range(H)]
        out["gl"] = [[pop.genotypeLikelihoods[i][g] for g in range(G)] for i in range(n_ind)]
        out["em"] = [[pop.EMLikelihoods[i][g] for g in range(G)] for i in range(n_ind)]
        out["gl_log_max"] = [pop.maxLogLikelihoods[i] for i in range(n_ind)]
        out["gof"] = [[pop.goodnessOfFitValues[g][i] for i in range(n_ind)] for g in range(G)]
        out["n_reads"] = [pop.nReads[i] for i in range(n_ind)]
        calls = []
        for gt in pop.genotypeCalls:
            calls.append(-1 if gt is None else [x is gt for x in genotypes].index(True))
        out["call"] = calls
        vars_out = []
        for key, v in uniq.items():
            # calculatePosterior with the flat prior 0.5 for every variant and, for substitutions, with the
            # reference's own prior (Variant.calculatePrior; the indel prior model of variant.pyx is prior
            # modelling outside this path and indexes Python-2 strings as char*, so it is not exercised)
            post_flat = pop.calculatePosterior(v, 1)
            prior = None
            post = None
            if len(key[1]) == len(key[2]):
                prior = v.calculatePrior(fa)
                post = pop.calculatePosterior(v, 0)
            holders = [h for h in range(H) if v in (<Haplotype>haps[h]).variants]
            vars_out.append((key[0], key[1], key[2], post_flat, prior, post, holders))
        out["variants"] = vars_out
    finally:
        pop = None
        buffers = []
        haps = []
        for (a, n) in all_arrays:
            _free_reads(<cAlignedRead**><size_t>a, n)
    return out


def population_seq(list hap_seqs, int win_start, int win_end, int hap_start, list per_ind_reads, int hla=0, int flank=0,
                   int use_em=0, int max_iters=100):
    """The same Population.setup + call, for haplotypes given as SEQUENCES (how the engine's batches carry them):
    haplotype h becomes a variant-free Haplotype over its own in-memory reference that holds exactly that sequence
    at [hap_start, hap_start + len), so Haplotype.__init__ cuts the very same bytes.  Needs
    win_start - hap_start = endBufferSize = min(2 * maxReadLength, 500) for some maxReadLength, i.e. an even flank
    <= 500, and len = (win_end - win_start) + 2 * flank.  Used by bench.py --impl reference to time the reference's
    own classes on the synthetic workload, and by tests to compare them with the oracle on it.
    Returns dict: freq, gl, em, gl_log_max, gof, call, ll (per haplotype, individual 0 only)."""
    cdef bytes name = b"chr"
    cdef int n_ind = len(per_ind_reads)
    cdef int H = len(hap_seqs)
    cdef int flank_len = win_start - hap_start
    if flank_len <= 0 or flank_len % 2 or flank_len > 500:
        raise ValueError("flank %d cannot be written as min(2 * maxReadLength, 500)" % flank_len)
    cdef int max_read_len = flank_len // 2
    # genomic coordinates only enter through differences: move the window next to the origin so that the
    # in-memory references stay as short as the haplotypes
    cdef int shift = hap_start - 8
    win_start -= shift
    win_end -= shift
    hap_start -= shift
    opts = _Options(flank, hla, n_ind, max(H, 2), use_em)
    cdef list keep = [], haps = [], buffers = [], all_arrays = []
    cdef Haplotype hap
    cdef bamReadBuffer buf
    cdef Population pop
    cdef cAlignedRead** arr
    cdef int i, g, h, k, n, G = H * (H + 1) // 2
    cdef bytes seq
    out = {}
    try:
        for seq in hap_seqs:
            if len(seq) != (win_end - win_start) + 2 * flank_len:
                raise ValueError("haplotype length %d != window + 2 * flank" % len(seq))
            # one spare base after the sequence: FastaFile.getSequence clamps endPos to seqLength - 1
            hap = Haplotype(name, win_start, win_end, (), MemFasta(name, b"N" * hap_start + seq + b"N"), max_read_len, opts)
            assert hap.hapLen == len(seq) and hap.cHaplotypeSequence[:hap.hapLen] == seq
            haps.append(hap)
        for (good, bad, broken) in per_ind_reads:
            buf = bamReadBuffer(name, win_start, win_end, opts)
            buf.sample = b"s"
            for lst, ra in ((good, buf.reads), (bad, buf.badReads), (broken, buf.brokenMates)):
                n = len(lst)
                arr = _make_reads(lst, keep, shift)
                all_arrays.append((<size_t>arr, n))
                for k in range(n):
                    (<cwindow.ReadArray>ra).append(arr[k])
                (<cwindow.ReadArray>ra).windowStart = (<cwindow.ReadArray>ra).array
                (<cwindow.ReadArray>ra).windowEnd = (<cwindow.ReadArray>ra).array + n
            buffers.append(buf)
        genotypes = generateAllGenotypesFromHaplotypeList(haps)
        pop = Population(opts)
        pop.setup([], haps, genotypes, n_ind, 0, buffers)
        pop.call(max_iters, 0)
        out["freq"] = [pop.frequencies[h] for h in range(H)]
        out["gl"] = [[pop.genotypeLikelihoods[i][g] for g in range(G)] for i in range(n_ind)]
        out["em"] = [[pop.EMLikelihoods[i][g] for g in range(G)] for i in range(n_ind)]
        out["gl_log_max"] = [pop.maxLogLikelihoods[i] for i in range(n_ind)]
        out["gof"] = [[pop.goodnessOfFitValues[g][i] for i in range(n_ind)] for g in range(G)]
        out["call"] = [-1 if gt is None else [x is gt for x in genotypes].index(True) for gt in pop.genotypeCalls]
    finally:
        pop = None
        buffers = []
        haps = []
        for (a, n) in all_arrays:
            _free_reads(<cAlignedRead**><size_t>a, n)
    return out


def select_haplotypes(bytes genome, int win_start, int win_end, list variants, list per_ind_good, int max_read_len=150,
                      int max_haplotypes=50, int original_max_haplotypes=50, int max_variants=8, int filter_by_coverage=1,
                      int coverage_sampling_level=30, int flank=0, list score_sets=None, list hap_score_sets=None):
    """One window through the reference's haplotype selection loop (src/cython/variantFilter.pyx:377-506
    getFilteredHaplotypes, :237-283 computeBestScoreForGenotype; excerpted into oracle/_ref/n1_ref by oracle/build.py).
    variants: [(refPos, removed, added, nSupportingReads)] in the window's order; per_ind_good: per individual the
    reads.windowStart..windowEnd list of (seq, qual, pos, end, mapq, bitFlag).
    Returns dict: selected = [tuple of variant indices per returned haplotype, in order], ref_seq (the reference
    haplotype's sequence), hap_seqs (Haplotype.cHaplotypeSequence of the returned haplotypes) and, when score_sets (a list
    of index tuples) is given, scores = computeBestScoreForGenotype(ref, Haplotype(set)) for each."""
    import n1_ref
    cdef bytes name = b"chr"
    cdef MemFasta fa = MemFasta(name, genome)
    cdef int n_ind = len(per_ind_good)
    opts = _Options(flank, 0, n_ind, max_haplotypes, 0)
    opts.rlen = max_read_len
    opts.maxHaplotypes = max_haplotypes
    opts.originalMaxHaplotypes = original_max_haplotypes
    opts.maxVariants = max_variants
    opts.filterVarsByCoverage = filter_by_coverage
    opts.coverageSamplingLevel = coverage_sampling_level
    cdef list keep = [], buffers = [], all_arrays = [], vobjs = []
    cdef bamReadBuffer buf
    cdef cAlignedRead** arr
    cdef Haplotype ref_hap, hap
    cdef int k, n
    out = {}
    try:
        for (p, rem, add, nsup) in variants:
            vobjs.append(Variant(name, p, rem, add, nsup, 1))
        for good in per_ind_good:
            buf = bamReadBuffer(name, win_start, win_end, opts)
            buf.sample = b"s"
            n = len(good)
            arr = _make_reads(good, keep)
            all_arrays.append((<size_t>arr, n))
            for k in range(n):
                buf.reads.append(arr[k])
            buf.reads.windowStart = buf.reads.array
            buf.reads.windowEnd = buf.reads.array + n
            for ra in (buf.badReads, buf.brokenMates):
                (<cwindow.ReadArray>ra).windowStart = (<cwindow.ReadArray>ra).array
                (<cwindow.ReadArray>ra).windowEnd = (<cwindow.ReadArray>ra).array
            buffers.append(buf)
        ref_hap = Haplotype(name, win_start, win_end, (), fa, max_read_len, opts)
        out["ref_seq"] = <bytes>ref_hap.cHaplotypeSequence[:ref_hap.hapLen]
        out["hap_start"] = ref_hap.startPos - ref_hap.endBufferSize
        sel = n1_ref.get_filtered_haplotypes(name, win_start, win_end, fa, opts, vobjs, ref_hap, buffers)
        index_of = {id(v): i for i, v in enumerate(vobjs)}
        out["selected"] = [tuple(index_of[id(v)] for v in vs) for vs in sel]
        seqs = []
        for vs in sel:
            hap = Haplotype(name, win_start, win_end, tuple(vs), fa, max_read_len, opts)
            seqs.append(<bytes>hap.cHaplotypeSequence[:hap.hapLen])
        out["hap_seqs"] = seqs
        if score_sets is not None:
            scores = []
            for idxs in score_sets:
                hap = Haplotype(name, win_start, win_end, tuple(vobjs[i] for i in idxs), fa, max_read_len, opts)
                scores.append(n1_ref.compute_best_score_for_genotype(buffers, ref_hap, hap, win_end - win_start,
                                                                     coverage_sampling_level))
            out["scores"] = scores
        if hap_score_sets is not None:   # computeBestScoreForHaplotype (variantFilter.pyx:212-234) of Haplotype(set)
            hscores = []
            for idxs in hap_score_sets:
                hap = Haplotype(name, win_start, win_end, tuple(vobjs[i] for i in idxs), fa, max_read_len, opts)
                hscores.append(n1_ref.compute_best_score_for_haplotype(buffers, hap))
            out["hap_scores"] = hscores
    finally:
        buffers = []
        for (a, n) in all_arrays:
            _free_reads(<cAlignedRead**><size_t>a, n)
    return out


def hla_haplotypes(bytes genome, int win_start, int win_end, list variants, list per_ind_good, int max_read_len=150,
                   int original_max_haplotypes=50, int coverage_sampling_level=30):
    """One window through the reference's --HLATyping haplotype selection (src/cython/variantFilter.pyx:655-736
    getAllHLAHaplotypesInRegion; excerpted into oracle/_ref/n1_ref by oracle/build.py).
    variants: [(refPos, removed, added, nSupportingReads, varSource)] in the window's order (varSource 2 = FILE_VAR: only
    those get a haplotype); per_ind_good as for select_haplotypes.
    Returns dict: haps = [variant index of every returned single-variant haplotype, in order - the list may repeat one],
    hap_scores / gt_scores = computeBestScoreForHaplotype of every FILE_VAR haplotype and computeBestScoreForGenotype of
    (best haplotype, it), hap_seqs = their sequences, ref_seq / hap_start as for select_haplotypes."""
    import n1_ref
    cdef bytes name = b"chr"
    cdef MemFasta fa = MemFasta(name, genome)
    cdef int n_ind = len(per_ind_good)
    opts = _Options(0, 1, n_ind, 50, 0)
    opts.rlen = max_read_len
    opts.maxHaplotypes = 50
    opts.originalMaxHaplotypes = original_max_haplotypes
    opts.coverageSamplingLevel = coverage_sampling_level
    cdef list keep = [], buffers = [], all_arrays = [], vobjs = []
    cdef bamReadBuffer buf
    cdef cAlignedRead** arr
    cdef Haplotype ref_hap, hap
    cdef int k, n
    out = {}
    try:
        for (p, rem, add, nsup, src) in variants:
            vobjs.append(Variant(name, p, rem, add, nsup, src))
        for good in per_ind_good:
            buf = bamReadBuffer(name, win_start, win_end, opts)
            buf.sample = b"s"
            n = len(good)
            arr = _make_reads(good, keep)
            all_arrays.append((<size_t>arr, n))
            for k in range(n):
                buf.reads.append(arr[k])
            buf.reads.windowStart = buf.reads.array
            buf.reads.windowEnd = buf.reads.array + n
            for ra in (buf.badReads, buf.brokenMates):
                (<cwindow.ReadArray>ra).windowStart = (<cwindow.ReadArray>ra).array
                (<cwindow.ReadArray>ra).windowEnd = (<cwindow.ReadArray>ra).array
            buffers.append(buf)
        ref_hap = Haplotype(name, win_start, win_end, (), fa, max_read_len, opts)
        out["ref_seq"] = <bytes>ref_hap.cHaplotypeSequence[:ref_hap.hapLen]
        out["hap_start"] = ref_hap.startPos - ref_hap.endBufferSize
        sel = n1_ref.get_all_hla_haplotypes(name, win_start, win_end, fa, opts, vobjs, ref_hap, buffers)
        index_of = {id(v): i for i, v in enumerate(vobjs)}
        out["haps"] = [index_of[id(vs[0])] for vs in sel]
        file_vars = [i for i, v in enumerate(variants) if v[4] == 2]
        haps = [Haplotype(name, win_start, win_end, (vobjs[i],), fa, max_read_len, opts) for i in file_vars]
        out["file_vars"] = file_vars
        out["hap_seqs"] = [<bytes>(<Haplotype>h).cHaplotypeSequence[:(<Haplotype>h).hapLen] for h in haps]
        hs = [n1_ref.compute_best_score_for_haplotype(buffers, h) for h in haps]
        out["hap_scores"] = hs
        if haps:
            best = max(zip(hs, haps))[1]            # the tuple order the reference's sorted(..., reverse=True)[0] uses
            out["gt_scores"] = [n1_ref.compute_best_score_for_genotype(buffers, best, h, win_end - win_start,
                                                                       coverage_sampling_level) for h in haps]
    finally:
        buffers = []
        for (a, n) in all_arrays:
            _free_reads(<cAlignedRead**><size_t>a, n)
    return out


def stage_reads(list reads, int region_start, int region_end, list windows, dict overrides=None):
    """Read staging (SURVEY 8f N3) through the reference's own bamReadBuffer: every read goes through
    addReadToBuffer -> checkAndTrimRead (src/cython/cwindow.pyx:560-595, 332-481), then
    ReadArray.setWindowPointers (cwindow.pyx:208-236) is asked for each window.

    reads      BAM records in file order as (seq, qual, [(cigar op, length), ...], chromID, pos, end, mapq, bitFlag,
               mateChromID, matePos, insertSize) - pos / end already as ReadIterator.get derives them
               (htslibWrapper.pyx:383-401; the decoder is the caller's)
    windows    [(start, end), ...]
    overrides  option values other than the defaults of runner.py:551-580
    Returns dict: good [n] (1 = reads list, 0 = badReads list), qual [n] (bytes after trimming), flag [n] (bitFlag after
    Read_SetQCFail), counts [7] (filteredReadCountsByType) and windows [(good_lo, good_hi, bad_lo, bad_hi)] as indices
    into the good / bad lists."""
    opts = _Options(0)
    for k, v in (overrides or {}).items():
        setattr(opts, k, v)
    cdef bytes name = b"chr"
    cdef bamReadBuffer buf = bamReadBuffer(name, region_start, region_end, opts)
    cdef int n = len(reads), i, j, nc
    cdef cAlignedRead** arr = <cAlignedRead**>calloc(n + 1, sizeof(cAlignedRead*))
    cdef cAlignedRead* r
    cdef bytes seq, qual
    cdef list keep = []
    cdef cwindow.ReadArray ga, ba
    out = {"good": [], "qual": [], "flag": [], "counts": [], "windows": []}
    try:
        for i in range(n):
            seq, qual, cigar, chrom_id, pos, end, mapq, flag, mate_chrom, mate_pos, isize = reads[i]
            keep.append(seq)
            r = <cAlignedRead*>calloc(1, sizeof(cAlignedRead))
            r.seq = <char*>seq
            r.rlen = len(seq)
            r.qual = <char*>malloc(r.rlen + 1)
            memcpy(r.qual, <char*>qual, r.rlen)
            r.qual[r.rlen] = 0
            nc = len(cigar)
            r.cigarLen = nc
            r.cigarOps = <short*>malloc(2 * max(nc, 1) * sizeof(short))
            for j in range(nc):
                r.cigarOps[2 * j] = cigar[j][0]
                r.cigarOps[2 * j + 1] = cigar[j][1]
            r.chromID = chrom_id
            r.pos = pos
            r.end = end
            r.mapq = mapq
            r.bitFlag = flag
            r.mateChromID = mate_chrom
            r.matePos = mate_pos
            r.insertSize = isize
            r.hash = NULL
            arr[i] = r
        for i in range(n):
            buf.addReadToBuffer(arr[i])
        ga = buf.reads
        ba = buf.badReads
        is_good = {}
        for i in range(ga.getSize()):
            is_good[<size_t>ga.array[i]] = 1
        for i in range(n):
            out["good"].append(1 if (<size_t>arr[i]) in is_good else 0)
            out["qual"].append(<bytes>arr[i].qual[:arr[i].rlen])
            out["flag"].append(arr[i].bitFlag)
        out["counts"] = [buf.filteredReadCountsByType[i] for i in range(7)]
        for (ws, we) in windows:
            ga.setWindowPointers(ws, we)
            ba.setWindowPointers(ws, we)
            out["windows"].append((ga.windowStart - ga.array, ga.windowEnd - ga.array,
                                   ba.windowStart - ba.array, ba.windowEnd - ba.array))
    finally:
        # the buffers' arrays hold borrowed pointers (destroyRead is a no-op in this build): free the reads here
        for i in range(n):
            if arr[i] != NULL:
                free(arr[i].qual)
                free(arr[i].cigarOps)
                free(arr[i])
        free(arr)
    return out
