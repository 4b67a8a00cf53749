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
    def __init__(self, flank):
        self.verbosity = 0
        self.calculateFlankScore = flank


cdef cAlignedRead** _make_reads(list reads, list keep) except NULL:
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
        r.pos = pos
        r.end = end
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
