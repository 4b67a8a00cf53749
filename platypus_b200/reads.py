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
import numpy as np  # noqa: F401  (array fields of the native staging path)
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


# ---- native staging: BAM record arrays -> the engine's packed read pool (plb_stage_reads_host) --------------------------

@dataclass
class BamRecords:
    """The alignment records of a BAM file as arrays of their own fields (PlbBamRecords, include/platypus_b200.h);
    sequences stay in BAM's 4-bit encoding."""
    ref_names: List[str]
    ref_id: "np.ndarray"
    pos: "np.ndarray"
    mapq: "np.ndarray"
    flag: "np.ndarray"
    mate_ref_id: "np.ndarray"
    mate_pos: "np.ndarray"
    tlen: "np.ndarray"
    cigar_off: "np.ndarray"
    cigar: "np.ndarray"
    seq_off: "np.ndarray"
    nib_off: "np.ndarray"
    nib: "np.ndarray"
    qual: "np.ndarray"

    @property
    def n(self):
        return len(self.pos)

    def as_struct(self):
        from . import _abi
        s = _abi.PlbBamRecords()
        s.n = self.n
        for k in ("ref_id", "pos", "mapq", "flag", "mate_ref_id", "mate_pos", "tlen", "cigar_off", "cigar", "seq_off",
                  "nib_off", "nib", "qual"):
            setattr(s, k, _abi.ptr(getattr(self, k)))
        return s


def read_bam_records(path, max_records=None) -> BamRecords:
    """Splits a BAM file into record-field arrays (BGZF = concatenated gzip members; record layout of the SAM
    specification 4.2).  No decoding: bases stay nibbles, CIGARs stay u32 words."""
    import numpy as np
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
    core, cig, nib, qual = [], [], [], []
    cig_off, seq_off, nib_off = [0], [0], [0]
    while off < len(data) and (max_records is None or len(core) < max_records):
        block_size, = struct.unpack_from("<i", data, off)
        rec = data[off + 4:off + 4 + block_size]
        off += 4 + block_size
        f = struct.unpack_from("<iiBBHHHiiii", rec, 0)
        core.append(f)
        _, _, l_name, _, _, n_cig, _, l_seq, _, _, _ = f
        p = 32 + l_name
        cig.append(rec[p:p + 4 * n_cig])
        p += 4 * n_cig
        nib.append(rec[p:p + (l_seq + 1) // 2])
        p += (l_seq + 1) // 2
        qual.append(rec[p:p + l_seq])
        cig_off.append(cig_off[-1] + n_cig)
        seq_off.append(seq_off[-1] + l_seq)
        nib_off.append(nib_off[-1] + (l_seq + 1) // 2)
    c = np.array(core, np.int64).reshape(-1, 11)
    return BamRecords(refs, c[:, 0].astype(np.int32), c[:, 1].astype(np.int32), c[:, 3].astype(np.uint8), c[:, 6].astype(np.uint16),
                      c[:, 8].astype(np.int32), c[:, 9].astype(np.int32), c[:, 10].astype(np.int32),
                      np.array(cig_off, np.int64), np.frombuffer(b"".join(cig) + b"\0\0\0\0", np.uint32).copy(),
                      np.array(seq_off, np.int64), np.array(nib_off, np.int64),
                      np.frombuffer(b"".join(nib) + b"\0", np.uint8).copy(), np.frombuffer(b"".join(qual) + b"\0", np.uint8).copy())


@dataclass
class StagedPool:
    """What plb_stage_reads_host leaves: per record the list it joined, its flag, pos / end and trimmed qualities, and the
    bases of all records as one 2-bit packed pool (+ exceptions) at the records' base offsets."""
    records: BamRecords
    kept: "np.ndarray"
    good: "np.ndarray"
    read_pos: "np.ndarray"
    read_end: "np.ndarray"
    flag: "np.ndarray"
    qual: "np.ndarray"
    seq2: "np.ndarray"
    exc_pos: "np.ndarray"
    exc_chr: "np.ndarray"
    counts: List[int]

    def good_index(self):
        import numpy as np
        return np.nonzero(self.kept & self.good)[0].astype(np.int32)

    def bad_index(self):
        import numpy as np
        return np.nonzero(self.kept & (1 - self.good))[0].astype(np.int32)

    def window_batch(self, windows, lib=None):
        """A PLB_SEQ_2BIT WindowBatch of one sample over this pool.  windows: [(start, end, hap_start, [haplotype bytes, ...])];
        each window gets the good and bad reads ReadArray.setWindowPointers selects (plb_window_slices_host).  The read
        arrays of the batch ARE the pool's arrays (no per-window copies; reads shared between windows are stored once)."""
        import ctypes as C
        import numpy as np
        from . import _abi
        from .batch import WindowBatch
        if lib is None:
            from .engine import load_library
            lib = load_library()
        W = len(windows)
        ws = np.array([w[0] for w in windows], np.int32)
        we = np.array([w[1] for w in windows], np.int32)
        lists = []
        for idx in (self.good_index(), self.bad_index()):
            pos, end = np.ascontiguousarray(self.read_pos[idx]), np.ascontiguousarray(self.read_end[idx])
            lo, hi = np.zeros(max(W, 1), np.int32), np.zeros(max(W, 1), np.int32)
            rc = lib.plb_window_slices_host(len(idx), _abi.ptr(pos), _abi.ptr(end), W, _abi.ptr(ws), _abi.ptr(we), _abi.ptr(lo), _abi.ptr(hi))
            if rc:
                raise RuntimeError(lib.plb_last_error().decode())
            lists.append((idx, lo, hi))
        (gi, glo, ghi), (bi, blo, bhi) = lists
        slots, off, n_good, n_bad = [], [0], [], []
        for w in range(W):
            g, b = gi[glo[w]:ghi[w]], bi[blo[w]:bhi[w]]
            slots += [g, b]
            n_good.append(len(g))
            n_bad.append(len(b))
            off.append(off[-1] + len(g) + len(b))
        hap_lens = [len(h) for w in windows for h in w[3]]
        hap_off = np.zeros(len(hap_lens) + 1, np.int64)
        np.cumsum(hap_lens, out=hap_off[1:])
        hap_ascii = np.frombuffer(b"".join(h for w in windows for h in w[3]) + b"\0", np.uint8)
        nh = int(hap_off[-1])
        hap2 = np.zeros((nh + 3) // 4 + 1, np.uint8)
        cap = max(16, int(np.count_nonzero(~np.isin(hap_ascii[:nh], np.frombuffer(b"ACGT", np.uint8)))))
        hpos, hchr, k = np.zeros(cap, np.int64), np.zeros(cap, np.uint8), C.c_int64(0)
        if lib.plb_pack_bases_host(_abi.ptr(hap_ascii), nh, _abi.ptr(hap2), 0, _abi.ptr(hpos), _abi.ptr(hchr), cap, C.byref(k)):
            raise RuntimeError(lib.plb_last_error().decode())
        win_hap_off = np.zeros(W + 1, np.int32)
        np.cumsum([len(w[3]) for w in windows], out=win_hap_off[1:])
        return WindowBatch(
            n_windows=W, n_individuals=1, win_hap_off=win_hap_off, win_start=ws, win_end=we,
            hap_start=np.array([w[2] for w in windows], np.int32), hap_seq_off=hap_off, hap_seq=hap2,
            wi_slot_off=np.array(off, np.int64), wi_n_good=np.array(n_good, np.int32), wi_n_bad=np.array(n_bad, np.int32),
            slot_read=np.concatenate(slots).astype(np.int32) if slots else np.zeros(0, np.int32),
            read_seq_off=self.records.seq_off, read_seq=self.seq2, read_qual=self.qual, read_pos=self.read_pos,
            read_end=self.read_end, read_mapq=self.records.mapq, read_qcfail=((self.flag & F_QCFAIL) != 0).astype(np.uint8),
            seq_format=_abi.PLB_SEQ_2BIT, read_exc_pos=self.exc_pos, read_exc_chr=self.exc_chr,
            hap_exc_pos=hpos[:k.value].copy(), hap_exc_chr=hchr[:k.value].copy())


def stage_records(records: BamRecords, options: Optional[ReadFilterOptions] = None, lib=None) -> StagedPool:
    """ReadIterator.get + addReadToBuffer + checkAndTrimRead for every record, and the packing of the bases, in the
    library (plb_stage_reads_host)."""
    import numpy as np
    from . import _abi
    if lib is None:
        from .engine import load_library
        lib = load_library()
    o = options or ReadFilterOptions()
    fo = _abi.PlbReadFilterOptions(o.min_good_qual_bases, o.min_map_qual, o.min_base_qual, o.trim_read_flank, o.trim_overlapping,
                                   o.trim_adapter, o.trim_soft_clipped, o.filter_duplicates, o.filter_mate_unmapped,
                                   o.filter_mate_distant, o.filter_small_insert)
    n, nb = records.n, int(records.seq_off[-1])
    kept, good = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.uint8)
    rpos, rend = np.zeros(max(n, 1), np.int32), np.zeros(max(n, 1), np.int32)
    flag, qual = np.zeros(max(n, 1), np.uint16), np.zeros(nb + 1, np.uint8)
    seq2 = np.zeros((nb + 3) // 4 + 2, np.uint8)
    cap = 1024
    while True:
        epos, echr = np.zeros(cap, np.int64), np.zeros(cap, np.uint8)
        out = _abi.PlbStagedReads(_abi.ptr(kept), _abi.ptr(good), _abi.ptr(rpos), _abi.ptr(rend), _abi.ptr(flag), _abi.ptr(qual),
                                  _abi.ptr(seq2), cap, 0, _abi.ptr(epos), _abi.ptr(echr))
        s = records.as_struct()
        rc = lib.plb_stage_reads_host(s, fo, out)
        if rc == _abi.PLB_ERR_SHAPE and cap < nb + 16:
            cap = min(nb + 16, cap * 16)      # more non-ACGT bases than expected: retry with room for them
            continue
        if rc:
            raise RuntimeError(lib.plb_last_error().decode())
        break
    k = int(out.n_exc)
    return StagedPool(records, kept[:n], good[:n], rpos[:n], rend[:n], flag[:n], qual, seq2, epos[:k].copy(), echr[:k].copy(),
                      [int(x) for x in out.counts])
