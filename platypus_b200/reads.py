"""
Host-side read staging: the step BEFORE the likelihood path.

The reference decodes BAM records through htslib into `cAlignedRead` structs
(src/cython/htslibWrapper.pyx:328-406), filters and quality-trims them while filling per-sample
buffers (src/cython/cwindow.pyx:332-481, 560-595) and bisects each buffer for the reads of a window
(src/cython/cwindow.pyx:208-236).  This module mirrors those three steps in plain Python / numpy so
that real alignments can be packed into a `WindowBatch` for the engine.  It is staging, not arithmetic
on the hot path: nothing here runs per (read, haplotype) pair.

BAM = BGZF = concatenated gzip members, so the standard library can inflate it; no htslib needed.
"""
import bisect
import gzip
import struct
from dataclasses import dataclass, field
from typing import List, Optional

from .batch import Read

BASE_LOOKUP = b"=ACMGRSVTWYHKDBN"                 # htslibWrapper.pyx:414-416

# SAM flag bits (htslibWrapper.pxd:233-296)
F_PAIRED, F_PROPER, F_UNMAPPED, F_MATE_UNMAPPED = 0x1, 0x2, 0x4, 0x8
F_REVERSE, F_MATE_REVERSE, F_SECONDARY, F_QCFAIL, F_DUPLICATE = 0x10, 0x20, 0x100, 0x200, 0x400


@dataclass
class AlignedRead:
    """The fields of cAlignedRead (src/cython/htslibWrapper.pxd:187-201)."""
    seq: bytes
    qual: bytearray          # raw phred; trimming sets entries to 0
    cigar: List[tuple]       # (op, length), op as in BAM: 0 M, 1 I, 2 D, 3 N, 4 S, 5 H, ...
    chrom_id: int
    pos: int                 # first base of the READ (soft clip at the start subtracted, .pyx:383-387)
    end: int                 # bam_endpos: one past the last reference base
    mapq: int
    flag: int
    mate_chrom_id: int
    mate_pos: int
    insert_size: int

    @property
    def rlen(self):
        return len(self.seq)

    def to_read(self) -> Read:
        return Read(self.seq, bytes(self.qual), self.pos, self.end, self.mapq, bool(self.flag & F_QCFAIL))


def decode_bam(path, max_records=None):
    """Yields (reference names, AlignedRead list) of a BAM file; the record layout is the BAM spec's,
    the field derivations follow ReadIterator.get (src/cython/htslibWrapper.pyx:328-406): records
    without sequence or without qualities (first byte 0xff) are dropped."""
    data = gzip.open(path, "rb").read()
    assert data[:4] == b"BAM\1", "not a BAM file"
    l_text, = struct.unpack_from("<i", data, 4)
    off = 8 + l_text
    n_ref, = struct.unpack_from("<i", data, off)
    off += 4
    refs = []
    for _ in range(n_ref):
        l_name, = struct.unpack_from("<i", data, off)
        refs.append(data[off + 4:off + 4 + l_name - 1].decode())
        off += 4 + l_name + 4
    reads = []
    while off < len(data) and (max_records is None or len(reads) < max_records):
        block_size, = struct.unpack_from("<i", data, off)
        rec = data[off + 4:off + 4 + block_size]
        off += 4 + block_size
        ref_id, pos, l_name, mapq, _bin, n_cig, flag, l_seq, m_ref, m_pos, tlen = struct.unpack_from("<iiBBHHHiiii", rec, 0)
        p = 32 + l_name
        cigar = []
        for k in range(n_cig):
            v, = struct.unpack_from("<I", rec, p + 4 * k)
            cigar.append((v & 0xF, v >> 4))
        p += 4 * n_cig
        packed = rec[p:p + (l_seq + 1) // 2]
        p += (l_seq + 1) // 2
        qual = rec[p:p + l_seq]
        if l_seq == 0 or qual[0] == 0xFF:
            continue
        seq = bytes(BASE_LOOKUP[(packed[i >> 1] >> (4 if (i & 1) == 0 else 0)) & 0xF] for i in range(l_seq))
        start = pos - (cigar[0][1] if cigar and cigar[0][0] == 4 else 0)
        ref_len = sum(n for op, n in cigar if op in (0, 2, 3, 7, 8))
        end = pos + (ref_len if ref_len > 0 else 1)          # bam_endpos
        reads.append(AlignedRead(seq, bytearray(qual), cigar, ref_id, start, end, mapq, flag, m_ref, m_pos, tlen))
    return refs, reads


@dataclass
class ReadFilterOptions:
    """Defaults of src/python/runner.py:551-580."""
    min_good_qual_bases: int = 20
    min_map_qual: int = 20
    min_base_qual: int = 20
    trim_read_flank: int = 0
    trim_overlapping: int = 1
    trim_adapter: int = 1
    trim_soft_clipped: int = 1
    filter_duplicates: int = 1
    filter_mate_unmapped: int = 1
    filter_mate_distant: int = 1
    filter_small_insert: int = 1


def check_and_trim_read(r: AlignedRead, last: Optional[AlignedRead], opt: ReadFilterOptions, counts: dict) -> bool:
    """checkAndTrimRead (src/cython/cwindow.pyx:332-481): False = the read goes to the bad-read list
    (most rejections also set the QC-fail flag, which makes the likelihood path skip the read);
    True = usable, with low-quality tails, overlapping mates, adapter read-through and soft clips
    set to quality 0."""
    def reject(kind=None, qcfail=True):
        if kind:
            counts[kind] = counts.get(kind, 0) + 1
        if qcfail:
            r.flag |= F_QCFAIL
        return False

    if r.flag & F_SECONDARY:
        return reject()
    if r.mapq < opt.min_map_qual:
        return reject("low_map_qual")
    n_low = sum(1 for q in r.qual if q < opt.min_base_qual)
    if r.rlen - n_low < opt.min_good_qual_bases:
        return reject("low_qual_bases")
    if r.flag & F_UNMAPPED:
        return reject("unmapped")
    paired = bool(r.flag & F_PAIRED)
    if opt.filter_mate_unmapped and paired and (r.flag & F_MATE_UNMAPPED):
        return reject("mate_unmapped", qcfail=False)
    if opt.filter_mate_distant and paired and (r.chrom_id != r.mate_chrom_id or not (r.flag & F_PROPER)):
        return reject("mate_distant", qcfail=False)
    if opt.filter_small_insert and paired and r.insert_size != 0 and abs(r.insert_size) < r.rlen:
        return reject("small_insert")
    if opt.filter_duplicates:
        if r.flag & F_DUPLICATE:
            return reject("duplicate")
        if last is not None and r.pos == last.pos and r.rlen == last.rlen:
            if paired:
                if last.mate_pos == r.mate_pos:
                    return reject("duplicate")
            else:
                return reject("duplicate")
    n = r.rlen
    if not (r.flag & F_REVERSE):                      # low-quality tail of a forward read
        for i in range(1, n + 1):
            if i < opt.trim_read_flank or r.qual[n - i] < 5:
                r.qual[n - i] = 0
            else:
                break
    else:
        for i in range(n):
            if i < opt.trim_read_flank or r.qual[i] < 5:
                r.qual[i] = 0
            else:
                break
    ins = abs(r.insert_size)
    if (opt.trim_overlapping == 1 and paired and ins > 0 and not (r.flag & F_REVERSE) and (r.flag & F_MATE_REVERSE)
            and ins < 2 * n):
        for i in range(1, min(n, (2 * n - r.insert_size) + 1) + 1):
            r.qual[n - i] = 0
    if opt.trim_adapter == 1 and paired and 0 < ins < n:
        if r.flag & F_REVERSE:
            for i in range(1, n - ins + 1):
                r.qual[n - i] = 0
        else:
            for i in range(ins, n):
                r.qual[i] = 0
    if opt.trim_soft_clipped == 1:
        idx = 0
        for op, ln in r.cigar:
            if op in (0, 1):
                idx += ln
            elif op == 4:
                for _ in range(ln):
                    r.qual[idx] = 0
                    idx += 1
    return True


@dataclass
class ReadBuffer:
    """bamReadBuffer (src/cython/cwindow.pyx:485-768) for one sample: good and bad reads in file order."""
    options: ReadFilterOptions = field(default_factory=ReadFilterOptions)
    reads: List[AlignedRead] = field(default_factory=list)
    bad_reads: List[AlignedRead] = field(default_factory=list)
    counts: dict = field(default_factory=dict)
    _last: Optional[AlignedRead] = None

    def add(self, r: AlignedRead):
        """addReadToBuffer (cwindow.pyx:560-595)."""
        ok = check_and_trim_read(r, self._last, self.options, self.counts)
        self._last = r
        (self.reads if ok else self.bad_reads).append(r)


def window_slice(reads: List[AlignedRead], start: int, end: int):
    """ReadArray.setWindowPointers (src/cython/cwindow.pyx:208-236) on a position-sorted list: the
    contiguous slice from the first read with pos >= max(1, start - longestRead) (skipping leading reads
    that end at or before `start`) up to the first read with pos >= end."""
    if not reads:
        return []
    longest = max(r.end - r.pos for r in reads)
    pos = [r.pos for r in reads]
    lo = bisect.bisect_left(pos, max(1, start - longest))
    hi = bisect.bisect_left(pos, end)
    while lo < len(reads) and reads[lo].end <= start:
        lo += 1
    return reads[lo:max(lo, min(hi, len(reads)))]
