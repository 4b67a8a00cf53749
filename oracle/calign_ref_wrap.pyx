# cython: language_level=3, boundscheck=False, wraparound=False
"""
Thin Python entry points onto the REFERENCE's own src/cython/calign.pyx (built unmodified
apart from the three non-algorithmic accommodations listed in oracle/build.py) so that
tests can pin the oracle's restatement of mapAndAlignReadToHaplotype against it.

TEST INFRASTRUCTURE ONLY.  This file contains no algorithm: it allocates the buffers the
reference expects (as chaplotype.pyx:190-191, 637-645 does) and forwards the call.
Only built when /root/reference is present; the result lives in oracle/_ref/.
"""
from libc.stdlib cimport malloc, free
from libc.string cimport memcpy
cimport calign
from htslibWrapper cimport cAlignedRead


def map_and_align(bytes read, bytes quals, int read_start, int hap_start, bytes hap, bytes gap_open,
                  int gap_extend=3, int nucprior=2, int hap_flank=1, int do_flank=0, int max_read_len=0,
                  bytes hash_read=None):
    """mapAndAlignReadToHaplotype (calign.pyx:170-272) for one (read, haplotype) pair.

    hash_read: when given, read.hash is built from THIS sequence instead of `read` - that is
    what alignReadToHaplotype does in HLA mode, where read/quals/readLen are clipped but
    read.hash stays that of the whole read (chaplotype.pyx:637-638, 647-655)."""
    cdef int read_len = len(read)
    cdef int hap_len = len(hap)
    cdef short* hh = NULL
    cdef short* hn = NULL
    cdef cAlignedRead r
    cdef int n_counts
    cdef int* counts
    cdef int score
    cdef char* c_read = read
    cdef char* c_quals = quals
    cdef char* c_hap = hap
    cdef char* c_go = gap_open
    if max_read_len < read_len:
        max_read_len = read_len
    if read_len < 7:
        return 0
    n_counts = 2 * (hap_len + max_read_len)
    counts = <int*>malloc(n_counts * sizeof(int))
    calign.hash_sequence_multihit(c_hap, hap_len, &hh, &hn)
    cdef char* c_hash_read = c_read
    cdef int hash_len = read_len
    if hash_read is not None:
        c_hash_read = hash_read
        hash_len = len(hash_read)
        if hash_len < read_len:
            raise ValueError("hash_read shorter than read")
    if hash_len > max_read_len:
        max_read_len = hash_len
    r.seq = c_hash_read
    r.qual = c_quals
    r.rlen = hash_len
    r.hash = NULL
    calign.hashReadForMapping(&r)
    score = calign.mapAndAlignReadToHaplotype(c_read, c_quals, read_start, hap_start, read_len, hap_len,
                                              hh, hn, r.hash, c_hap, gap_extend, nucprior, c_go,
                                              counts, n_counts, hap_flank, do_flank)
    free(r.hash)
    free(hh)
    free(hn)
    free(counts)
    return score


def hap_hash_table(bytes hap):
    """hash_sequence_multihit (calign.pyx:94-124): returns (head list, next list)."""
    cdef short* hh = NULL
    cdef short* hn = NULL
    cdef char* c_hap = hap
    cdef int i
    calign.hash_sequence_multihit(c_hap, len(hap), &hh, &hn)
    head = [hh[i] for i in range(16384)]
    nxt = [hn[i] for i in range(16384)]
    free(hh)
    free(hn)
    return head, nxt


def read_hashes(bytes read):
    """hashReadForMapping (calign.pyx:155-165)."""
    cdef cAlignedRead r
    cdef char* c_read = read
    cdef int i
    cdef int n = len(read)
    if n < 8:
        return []
    r.seq = c_read
    r.rlen = n
    r.hash = NULL
    calign.hashReadForMapping(&r)
    out = [r.hash[i] for i in range(n - 7)]
    free(r.hash)
    return out
