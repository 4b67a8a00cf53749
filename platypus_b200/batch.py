"""
Host-side container for a batch of windows: the flat (CSR) equivalent of what
callVariantsInWindow hands to Population.setup one window at a time
(reference: src/cython/variantcaller.pyx:74-141, src/cython/cpopulation.pyx:197-309).

The arrays are exactly the fields of PlbWindowBatch in include/platypus_b200.h.
"""
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi


@dataclass
class Read:
    """The fields of cAlignedRead the likelihood path touches
    (reference: src/cython/htslibWrapper.pxd:187-201)."""
    seq: bytes
    qual: bytes          # raw phred, not +33
    pos: int
    end: int
    mapq: int = 60
    qcfail: bool = False

    @property
    def rlen(self):
        return len(self.seq)


@dataclass
class Window:
    """One window: haplotype sequences sharing an interval + per-individual read lists."""
    start: int                       # Haplotype.startPos
    end: int                         # Haplotype.endPos
    hap_start: int                   # startPos - endBufferSize (chaplotype.pyx:604)
    haplotypes: List[bytes]
    # per individual: (good, bad, broken) lists of Read
    reads: List[Sequence[Sequence[Read]]]
    hap_var_mask: Optional[List[int]] = None   # per haplotype bit mask of contained variants
    var_prior: Optional[List[float]] = None    # per variant prior


@dataclass
class WindowBatch:
    n_windows: int
    n_individuals: int
    win_hap_off: np.ndarray
    win_start: np.ndarray
    win_end: np.ndarray
    hap_start: np.ndarray
    hap_seq_off: np.ndarray
    hap_seq: np.ndarray
    wi_slot_off: np.ndarray
    wi_n_good: np.ndarray
    wi_n_bad: np.ndarray
    slot_read: np.ndarray
    read_seq_off: np.ndarray
    read_seq: np.ndarray
    read_qual: np.ndarray
    read_pos: np.ndarray
    read_end: np.ndarray
    read_mapq: np.ndarray
    read_qcfail: np.ndarray
    max_variants: int = 0
    win_n_var: Optional[np.ndarray] = None
    hap_var_mask: Optional[np.ndarray] = None
    var_prior: Optional[np.ndarray] = None
    # 2-bit packed bases (PLB_SEQ_2BIT, include/platypus_b200.h): hap_seq / read_seq then hold 4 bases per byte and the
    # bases outside ACGT travel in the exception lists; offsets keep counting bases
    seq_format: int = 0
    read_exc_pos: Optional[np.ndarray] = None
    read_exc_chr: Optional[np.ndarray] = None
    hap_exc_pos: Optional[np.ndarray] = None
    hap_exc_chr: Optional[np.ndarray] = None
    # packed qualities (PlbWindowBatch.qual_bits): read_qual then holds 4- or 6-bit codes into qual_table
    qual_bits: int = 0
    qual_table: Optional[np.ndarray] = None
    _keep: list = field(default_factory=list, repr=False)

    # ---- construction -------------------------------------------------------------------
    @classmethod
    def from_windows(cls, windows: Sequence[Window], n_individuals: int, dedupe_reads: bool = True):
        """Pack Window objects.  Read objects shared between windows (same Python object) are
        stored once in the pool, like the reference's shared cAlignedRead pointers."""
        W = len(windows)
        win_hap_off = np.zeros(W + 1, np.int32)
        hap_chunks, hap_lens = [], []
        pool, pool_ix = [], {}
        slot_read, wi_off, n_good, n_bad = [], [0], [], []
        max_var = 0
        for w, win in enumerate(windows):
            assert len(win.reads) == n_individuals, "every window needs one read-list triple per individual"
            for h in win.haplotypes:
                hap_chunks.append(np.frombuffer(h, np.uint8))
                hap_lens.append(len(h))
            win_hap_off[w + 1] = win_hap_off[w] + len(win.haplotypes)
            for i in range(n_individuals):
                good, bad, broken = win.reads[i]
                for r in list(good) + list(bad) + list(broken):
                    key = id(r) if dedupe_reads else len(pool)
                    j = pool_ix.get(key)
                    if j is None:
                        j = len(pool)
                        pool_ix[key] = j
                        pool.append(r)
                    slot_read.append(j)
                n_good.append(len(good))
                n_bad.append(len(bad))
                wi_off.append(len(slot_read))
            if win.var_prior is not None:
                max_var = max(max_var, len(win.var_prior))
        hap_seq_off = np.zeros(len(hap_lens) + 1, np.int64)
        np.cumsum(hap_lens, out=hap_seq_off[1:])
        read_lens = [len(r.seq) for r in pool]
        read_seq_off = np.zeros(len(pool) + 1, np.int64)
        np.cumsum(read_lens, out=read_seq_off[1:])
        for r in pool:
            assert len(r.seq) == len(r.qual), "seq/qual length mismatch"
        b = cls(
            n_windows=W, n_individuals=n_individuals,
            win_hap_off=win_hap_off,
            win_start=np.array([w.start for w in windows], np.int32),
            win_end=np.array([w.end for w in windows], np.int32),
            hap_start=np.array([w.hap_start for w in windows], np.int32),
            hap_seq_off=hap_seq_off,
            hap_seq=np.concatenate(hap_chunks) if hap_chunks else np.zeros(0, np.uint8),
            wi_slot_off=np.array(wi_off, np.int64),
            wi_n_good=np.array(n_good, np.int32), wi_n_bad=np.array(n_bad, np.int32),
            slot_read=np.array(slot_read, np.int32),
            read_seq_off=read_seq_off,
            read_seq=np.frombuffer(b"".join(r.seq for r in pool), np.uint8).copy() if pool else np.zeros(0, np.uint8),
            read_qual=np.frombuffer(b"".join(r.qual for r in pool), np.uint8).copy() if pool else np.zeros(0, np.uint8),
            read_pos=np.array([r.pos for r in pool], np.int32),
            read_end=np.array([r.end for r in pool], np.int32),
            read_mapq=np.array([r.mapq for r in pool], np.uint8),
            read_qcfail=np.array([1 if r.qcfail else 0 for r in pool], np.uint8),
        )
        if max_var > 0:
            b.max_variants = max_var
            b.win_n_var = np.array([len(w.var_prior) if w.var_prior is not None else 0 for w in windows], np.int32)
            masks = []
            for w in windows:
                masks += list(w.hap_var_mask) if w.hap_var_mask is not None else [0] * len(w.haplotypes)
            b.hap_var_mask = np.array(masks, np.uint64)
            pri = np.zeros((W, max_var), np.float64)
            for i, w in enumerate(windows):
                if w.var_prior is not None:
                    pri[i, :len(w.var_prior)] = w.var_prior
            b.var_prior = pri
        return b

    # ---- derived sizes ------------------------------------------------------------------
    @property
    def n_haps(self):
        return int(self.win_hap_off[-1])

    @property
    def n_reads(self):
        return int(len(self.read_pos))

    @property
    def n_slots(self):
        return int(self.wi_slot_off[-1])

    def haps_per_window(self):
        return np.diff(self.win_hap_off)

    def max_haps(self):
        return int(self.haps_per_window().max()) if self.n_windows else 0

    def ll_offsets(self):
        """[W*nInd+1] offsets of the per-(window,individual) [H][T] blocks of PlbLoglikOut."""
        H = np.repeat(self.haps_per_window().astype(np.int64), self.n_individuals)
        T = np.diff(self.wi_slot_off)
        off = np.zeros(self.n_windows * self.n_individuals + 1, np.int64)
        np.cumsum(H * T, out=off[1:])
        return off

    def input_nbytes(self):
        return sum(int(a.nbytes) for a in self._arrays() if a is not None)

    def _arrays(self):
        return [self.win_hap_off, self.win_start, self.win_end, self.hap_start, self.hap_seq_off, self.hap_seq,
                self.wi_slot_off, self.wi_n_good, self.wi_n_bad, self.slot_read, self.read_seq_off, self.read_seq,
                self.read_qual, self.read_pos, self.read_end, self.read_mapq, self.read_qcfail, self.win_n_var,
                self.hap_var_mask, self.var_prior, self.read_exc_pos, self.read_exc_chr, self.hap_exc_pos, self.hap_exc_chr]

    def n_bases(self):
        """(haplotype bases, read bases) of the batch, whatever the sequence format."""
        return int(self.hap_seq_off[-1]), int(self.read_seq_off[-1])

    def pack(self, lib=None, quals=True) -> "WindowBatch":
        """The same batch with 2-bit packed bases (PLB_SEQ_2BIT) and, with `quals`, 4- / 6-bit packed qualities: what
        the staging step hands to the GPU instead of ASCII (BAM's 4-bit nibbles pack straight into it,
        htslibWrapper.pyx:414-416).  Packing is done by the library's host helpers; no GPU work."""
        import ctypes as C
        import dataclasses
        if self.seq_format == _abi.PLB_SEQ_2BIT:
            return self.pack_quals(lib) if quals else self
        if lib is None:
            from .engine import load_library
            lib = load_library()

        def pack(seq, n):
            src = np.ascontiguousarray(seq[:n], np.uint8)
            dst = np.zeros((n + 3) // 4 + 1, np.uint8)
            cap = max(16, int(np.count_nonzero(~np.isin(src, np.frombuffer(b"ACGT", np.uint8)))))
            pos, chr_ = np.zeros(cap, np.int64), np.zeros(cap, np.uint8)
            k = C.c_int64(0)
            rc = lib.plb_pack_bases_host(_abi.ptr(src), n, _abi.ptr(dst), 0, _abi.ptr(pos), _abi.ptr(chr_), cap, C.byref(k))
            if rc:
                raise RuntimeError("plb_pack_bases_host failed: %s" % lib.plb_last_error().decode())
            return dst, pos[:k.value].copy(), chr_[:k.value].copy()
        nh, nr = self.n_bases()
        hs, hp, hc = pack(self.hap_seq, nh)
        rs, rp, rc_ = pack(self.read_seq, nr)
        out = dataclasses.replace(self, hap_seq=hs, read_seq=rs, seq_format=_abi.PLB_SEQ_2BIT, read_exc_pos=rp,
                                  read_exc_chr=rc_, hap_exc_pos=hp, hap_exc_chr=hc, _keep=[])
        return out.pack_quals(lib) if quals else out

    def pack_quals(self, lib=None) -> "WindowBatch":
        """The same batch with the base qualities as 4- or 6-bit codes into a table of the batch's distinct values
        (plb_pack_quals_host; lossless).  A batch with more than 64 distinct qualities is returned unchanged."""
        import ctypes as C
        import dataclasses
        if self.qual_bits:
            return self
        if lib is None:
            from .engine import load_library
            lib = load_library()
        nr = int(self.read_seq_off[-1])
        src = np.ascontiguousarray(self.read_qual[:nr], np.uint8)
        dst = np.zeros((nr * 6 + 7) // 8 + 8, np.uint8)
        table, bits = np.zeros(64, np.uint8), C.c_int32(0)
        rc = lib.plb_pack_quals_host(_abi.ptr(src), nr, _abi.ptr(dst), C.byref(bits), _abi.ptr(table))
        if rc == _abi.PLB_ERR_SHAPE:
            return self
        if rc:
            raise RuntimeError("plb_pack_quals_host failed: %s" % lib.plb_last_error().decode())
        return dataclasses.replace(self, read_qual=dst[:(nr * bits.value + 7) // 8 + 4].copy(), qual_bits=bits.value, qual_table=table,
                                   _keep=[])

    # ---- ABI ----------------------------------------------------------------------------
    def as_struct(self):
        """PlbWindowBatch with HOST pointers into this object's arrays (keep `self` alive)."""
        def c(a, dt):
            if a is None:
                return None
            a2 = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a2)
            return a2
        s = _abi.PlbWindowBatch()
        s.n_windows, s.n_individuals = self.n_windows, self.n_individuals
        s.n_haps, s.n_reads, s.n_slots = self.n_haps, self.n_reads, self.n_slots
        self._keep.clear()
        for name, dt in (("win_hap_off", np.int32), ("win_start", np.int32), ("win_end", np.int32),
                         ("hap_start", np.int32), ("hap_seq_off", np.int64), ("hap_seq", np.uint8),
                         ("wi_slot_off", np.int64), ("wi_n_good", np.int32), ("wi_n_bad", np.int32),
                         ("slot_read", np.int32), ("read_seq_off", np.int64), ("read_seq", np.uint8),
                         ("read_qual", np.uint8), ("read_pos", np.int32), ("read_end", np.int32),
                         ("read_mapq", np.uint8), ("read_qcfail", np.uint8), ("win_n_var", np.int32),
                         ("hap_var_mask", np.uint64), ("var_prior", np.float64)):
            setattr(s, name, _abi.ptr(c(getattr(self, name), dt)))
        s.max_variants = int(self.max_variants)
        s.seq_format = int(self.seq_format)
        s.qual_bits = int(self.qual_bits)
        if self.qual_bits:
            C_ = __import__("ctypes")
            C_.memmove(s.qual_table, np.ascontiguousarray(self.qual_table, np.uint8).ctypes.data, 64)
        if self.seq_format:
            s.n_read_exc = 0 if self.read_exc_pos is None else len(self.read_exc_pos)
            s.n_hap_exc = 0 if self.hap_exc_pos is None else len(self.hap_exc_pos)
            for name, dt in (("read_exc_pos", np.int64), ("read_exc_chr", np.uint8), ("hap_exc_pos", np.int64),
                             ("hap_exc_chr", np.uint8)):
                setattr(s, name, _abi.ptr(c(getattr(self, name), dt)))
        return s

    # ---- sharding (contiguous blocks of windows per GPU) ---------------------
    def slice_windows(self, lo: int, hi: int) -> "WindowBatch":
        """Sub-batch with windows [lo, hi); the read pool is re-indexed to the reads the
        shard touches (reads straddling a shard boundary are duplicated into both shards, as
        the reference duplicates them across regions)."""
        assert self.seq_format == 0 and self.qual_bits == 0, "slice the ASCII batch, then pack()"
        nI = self.n_individuals
        h0, h1 = int(self.win_hap_off[lo]), int(self.win_hap_off[hi])
        s0, s1 = int(self.wi_slot_off[lo * nI]), int(self.wi_slot_off[hi * nI])
        slots = self.slot_read[s0:s1]
        used, inv = np.unique(slots, return_inverse=True)
        rlen = np.diff(self.read_seq_off)[used]
        roff = np.zeros(len(used) + 1, np.int64)
        np.cumsum(rlen, out=roff[1:])
        idx = _gather_index(self.read_seq_off, used) if len(used) else np.zeros(0, np.int64)
        hb0, hb1 = int(self.hap_seq_off[h0]), int(self.hap_seq_off[h1])
        return WindowBatch(
            n_windows=hi - lo, n_individuals=nI,
            win_hap_off=(self.win_hap_off[lo:hi + 1] - h0).astype(np.int32),
            win_start=self.win_start[lo:hi].copy(), win_end=self.win_end[lo:hi].copy(),
            hap_start=self.hap_start[lo:hi].copy(),
            hap_seq_off=(self.hap_seq_off[h0:h1 + 1] - hb0).astype(np.int64),
            hap_seq=self.hap_seq[hb0:hb1].copy(),
            wi_slot_off=(self.wi_slot_off[lo * nI:hi * nI + 1] - s0).astype(np.int64),
            wi_n_good=self.wi_n_good[lo * nI:hi * nI].copy(), wi_n_bad=self.wi_n_bad[lo * nI:hi * nI].copy(),
            slot_read=inv.astype(np.int32),
            read_seq_off=roff, read_seq=self.read_seq[idx], read_qual=self.read_qual[idx],
            read_pos=self.read_pos[used], read_end=self.read_end[used], read_mapq=self.read_mapq[used],
            read_qcfail=self.read_qcfail[used],
            max_variants=self.max_variants,
            win_n_var=None if self.win_n_var is None else self.win_n_var[lo:hi].copy(),
            hap_var_mask=None if self.hap_var_mask is None else self.hap_var_mask[h0:h1].copy(),
            var_prior=None if self.var_prior is None else self.var_prior[lo:hi].copy(),
        )


def with_haplotypes(ref_batch: "WindowBatch", win_hap_off, hap_seqs, hap_masks, variants=None, var_prior=None) -> "WindowBatch":
    """The batch of the window model after haplotype selection: the windows, slots and reads of `ref_batch`, with the
    haplotype arrays replaced by `hap_seqs` (list of bytes in window order, the reference haplotype first - what
    variantcaller.pyx:116-120 hands to Population.setup) and `hap_masks` (bit v = window variant v) as hap_var_mask."""
    import dataclasses
    lens = np.fromiter((len(h) for h in hap_seqs), np.int64, len(hap_seqs))
    off = np.zeros(len(hap_seqs) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    W = ref_batch.n_windows
    b = dataclasses.replace(
        ref_batch, win_hap_off=np.asarray(win_hap_off, np.int32), hap_seq_off=off,
        hap_seq=np.frombuffer(b"".join(hap_seqs) + b"\0", np.uint8).copy(), _keep=[])
    if variants is not None:
        n_var = np.diff(variants.win_var_off).astype(np.int32)
        mv = max(1, int(n_var.max()) if W else 1)
        b.max_variants = mv
        b.win_n_var = n_var
        b.hap_var_mask = np.asarray(hap_masks, np.uint64)
        b.var_prior = np.full((W, mv), 0.5) if var_prior is None else np.asarray(var_prior, np.float64)
    return b


def concat_batches(parts: Sequence["WindowBatch"]) -> "WindowBatch":
    """Concatenate batches window-wise (read pools are concatenated and re-indexed)."""
    parts = [p for p in parts if p.n_windows > 0]
    assert parts, "nothing to concatenate"
    nI = parts[0].n_individuals
    assert all(p.n_individuals == nI for p in parts)

    def cat_off(name, dt):
        out, base = [np.zeros(1, dt)], 0
        for p in parts:
            a = getattr(p, name)
            out.append(a[1:].astype(dt) + base)
            base += int(a[-1])
        return np.concatenate(out)

    def cat(name):
        return np.concatenate([getattr(p, name) for p in parts])

    read_base = np.cumsum([0] + [p.n_reads for p in parts])
    b = WindowBatch(
        n_windows=sum(p.n_windows for p in parts), n_individuals=nI,
        win_hap_off=cat_off("win_hap_off", np.int32), win_start=cat("win_start"), win_end=cat("win_end"),
        hap_start=cat("hap_start"), hap_seq_off=cat_off("hap_seq_off", np.int64), hap_seq=cat("hap_seq"),
        wi_slot_off=cat_off("wi_slot_off", np.int64), wi_n_good=cat("wi_n_good"), wi_n_bad=cat("wi_n_bad"),
        slot_read=np.concatenate([p.slot_read + np.int32(read_base[i]) for i, p in enumerate(parts)]).astype(np.int32),
        read_seq_off=cat_off("read_seq_off", np.int64), read_seq=cat("read_seq"), read_qual=cat("read_qual"),
        read_pos=cat("read_pos"), read_end=cat("read_end"), read_mapq=cat("read_mapq"), read_qcfail=cat("read_qcfail"),
    )
    if all(p.win_n_var is not None for p in parts):
        mv = max(p.max_variants for p in parts)
        b.max_variants = mv
        b.win_n_var = cat("win_n_var")
        b.hap_var_mask = cat("hap_var_mask")
        pri = np.zeros((b.n_windows, mv), np.float64)
        o = 0
        for p in parts:
            pri[o:o + p.n_windows, :p.max_variants] = p.var_prior
            o += p.n_windows
        b.var_prior = pri
    return b


def _gather_index(off, used):
    """Vectorised concatenation of ranges [off[r], off[r+1]) for r in used."""
    lens = (off[used + 1] - off[used]).astype(np.int64)
    total = int(lens.sum())
    starts = np.repeat(off[used], lens)
    within = np.arange(total, dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
    return starts + within


def shard_bounds(n_windows: int, world_size: int):
    """Contiguous, near-equal window blocks per rank."""
    base, rem = divmod(n_windows, world_size)
    bounds = [0]
    for r in range(world_size):
        bounds.append(bounds[-1] + base + (1 if r < rem else 0))
    return bounds


# ---- sites (scope row N4: the per-sample loop of outputCallToVCF) --------------------------------
@dataclass
class SiteBatch:
    """Reported positions of a WindowBatch and what computeGenotypeCallAndLikelihoods needs per site
    (reference: src/cython/vcfutils.pyx:391-417): the window, the window-local indices of the variants
    reported at the site (variantsThisPos order) and haplotypeIsRefAtThisPos per haplotype."""
    site_win: np.ndarray        # int32 [S]
    site_var_off: np.ndarray    # int32 [S+1]
    site_var: np.ndarray        # int32
    site_hap_off: np.ndarray    # int64 [S+1]
    hap_is_ref: np.ndarray      # uint8
    min_posterior: int = 5

    @property
    def n_sites(self):
        return len(self.site_win)

    def max_pairs(self):
        nv = int(np.max(np.diff(self.site_var_off))) if self.n_sites else 1
        return (nv + 1) * (nv + 2) // 2

    def as_struct(self):
        from . import _abi
        s = _abi.PlbSiteBatch()
        s.n_sites = self.n_sites
        for name in ("site_win", "site_var_off", "site_var", "site_hap_off", "hap_is_ref"):
            setattr(s, name, _abi.ptr(getattr(self, name)))
        s.min_posterior = int(self.min_posterior)
        return s

    @staticmethod
    def from_lists(batch, sites, min_posterior=5):
        """sites: list of (window, [variant indices], [is_ref per haplotype of the window])."""
        sw, svo, sv, sho, ref = [], [0], [], [0], []
        for w, vs, isref in sites:
            H = int(batch.win_hap_off[w + 1] - batch.win_hap_off[w])
            assert len(isref) == H
            sw.append(w)
            sv.extend(vs)
            svo.append(len(sv))
            ref.extend(isref)
            sho.append(len(ref))
        return SiteBatch(np.asarray(sw, np.int32), np.asarray(svo, np.int32),
                         np.asarray(sv, np.int32).reshape(-1), np.asarray(sho, np.int64),
                         np.asarray(ref, np.uint8).reshape(-1), min_posterior)


def alloc_site_out(batch, sites):
    S, nI, Pm = sites.n_sites, batch.n_individuals, sites.max_pairs()
    return {"max_pairs": Pm, "phased": np.zeros((S, nI, 2), np.int32), "lik": np.zeros((S, nI, Pm)),
            "post": np.zeros((S, nI, 3)), "phred": np.zeros((S, nI, 3), np.int32), "gof": np.zeros((S, nI)),
            "gt": np.zeros((S, nI, 2), np.int32), "gl_log10": np.zeros((S, nI, 3))}


def site_out_struct(arrs):
    from . import _abi
    o = _abi.PlbSiteOut()
    o.max_pairs = arrs["max_pairs"]
    for k in ("phased", "lik", "post", "phred", "gof", "gt", "gl_log10"):
        setattr(o, k, _abi.ptr(arrs[k]))
    return o


# ---- N1: candidate variants of a batch of windows ------------------------------------------------------------------

@dataclass
class VariantSet:
    """The candidate variants of every window of a batch (PlbVariantSet in include/platypus_b200.h): the fields of the
    reference's Variant that haplotype construction and the selection loop read (src/cython/variant.pyx:109-145).
    Per window in the order of the window's sorted `variants` list."""
    win_var_off: np.ndarray
    var_pos: np.ndarray
    var_n_removed: np.ndarray
    var_n_support: np.ndarray
    var_added_off: np.ndarray
    var_added: np.ndarray

    @classmethod
    def from_lists(cls, per_window):
        """per_window: for every window a list of (refPos, removed | len(removed), added, nSupportingReads)."""
        off, pos, nrem, nsup, aoff, added = [0], [], [], [], [0], bytearray()
        for vs in per_window:
            for v in vs:
                p, rem, add = v[0], v[1], v[2]
                pos.append(p)
                nrem.append(rem if isinstance(rem, int) else len(rem))
                nsup.append(v[3] if len(v) > 3 else 1)
                added += add
                aoff.append(len(added))
            off.append(len(pos))
        return cls(np.asarray(off, np.int32), np.asarray(pos, np.int32), np.asarray(nrem, np.int32),
                   np.asarray(nsup, np.int32), np.asarray(aoff, np.int64),
                   np.frombuffer(bytes(added) + b"\0", np.uint8).copy())

    def n_vars(self, w):
        return int(self.win_var_off[w + 1] - self.win_var_off[w])

    def slice_windows(self, lo: int, hi: int) -> "VariantSet":
        """The variants of windows [lo, hi) (offsets rebased)."""
        v0, v1 = int(self.win_var_off[lo]), int(self.win_var_off[hi])
        a0, a1 = int(self.var_added_off[v0]), int(self.var_added_off[v1])
        return VariantSet((self.win_var_off[lo:hi + 1] - v0).astype(np.int32), self.var_pos[v0:v1].copy(),
                          self.var_n_removed[v0:v1].copy(), self.var_n_support[v0:v1].copy(),
                          (self.var_added_off[v0:v1 + 1] - a0).astype(np.int64),
                          np.concatenate([self.var_added[a0:a1], np.zeros(1, np.uint8)]))

    def as_struct(self):
        s = _abi.PlbVariantSet()
        for k in ("win_var_off", "var_pos", "var_n_removed", "var_n_support", "var_added_off", "var_added"):
            setattr(s, k, _abi.ptr(getattr(self, k)))
        return s
