"""
ctypes binding of libplatypus_b200.so (include/platypus_b200.h) and a thin Engine object.

There is no CPU fallback: if the shared library is missing, or no B200-class GPU is visible,
construction raises.  PyTorch is used only as plumbing (device memory for the device-resident
path, streams, torch.distributed); all arithmetic happens in the CUDA kernels.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from .batch import WindowBatch

_HERE = os.path.dirname(os.path.abspath(__file__))
# PLB_LIBRARY: another build of the same CUDA library (A/B experiments: tools build variants next to the default one)
LIB_PATH = os.environ.get("PLB_LIBRARY") or os.path.join(_HERE, "libplatypus_b200.so")
_lib = None


class PlbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("platypus_b200 error %d: %s" % (code, msg))
        self.code = code


def load_library():
    """dlopen the in-tree CUDA library; raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PlbError(_abi.PLB_ERR_CUDA, "%s not built (run python -c 'import __graft_entry__ as g; g.build()'); "
                           "there is no CPU fallback" % LIB_PATH)
        _lib = _abi.declare(C.CDLL(LIB_PATH))
    return _lib


def _check(lib, rc):
    if rc < 0:
        raise PlbError(rc, lib.plb_last_error().decode("utf-8", "replace"))
    return rc


class Engine:
    """One engine per GPU.  Mirrors the three seams of the reference:
    S1 fast_align/align_batch, S2 window_loglik, S3 population_run."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = C.c_void_p()
        _check(self.lib, self.lib.plb_context_create(int(device), C.c_void_p(stream) if stream else None, C.byref(h)))
        self.ctx = h
        self.device = device

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.plb_context_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self):
        return int(self.lib.plb_launch_count(self.ctx))

    def last_stats(self):
        st = _abi.PlbRunStats()
        _check(self.lib, self.lib.plb_last_stats(self.ctx, C.byref(st)))
        return st.as_dict()

    def set_timing(self, on=True):
        _check(self.lib, self.lib.plb_set_timing(self.ctx, 1 if on else 0))

    def kernel_times(self):
        """Mean device time (ms) of each kernel over the runs since set_timing(True), from CUDA events
        on the context stream.  Returns (dict, number of runs averaged)."""
        ms = (C.c_float * 6)()
        n = _check(self.lib, self.lib.plb_kernel_times(self.ctx, ms))
        return dict(zip(_abi.KERNEL_NAMES, [float(x) for x in ms])), n

    # ---- S1 ---------------------------------------------------------------------------------
    def fast_align(self, seq1: bytes, seq2: bytes, qual2: bytes, gap_open: bytes, gapextend=3, nucprior=2):
        """fastAlignmentRoutine(seq1, seq2, qual2, len1, len2, gapextend, nucprior, localgapopen, NULL, NULL, NULL)
        (reference: src/c/align.h:8-9)."""
        L = len(seq2)
        if len(seq1) < L + 15 or len(gap_open) < L + 15 or len(qual2) != L:
            raise PlbError(_abi.PLB_ERR_SHAPE, "need len(seq1) >= len(seq2)+15 and matching qual/gap-open lengths")
        return _check(self.lib, self.lib.plb_fast_align(self.ctx, seq1, seq2, qual2, L + 15, L, gapextend, nucprior,
                                                        gap_open, None, None, None))

    def fast_align_traceback(self, seq1: bytes, seq2: bytes, qual2: bytes, gap_open: bytes, gapextend=3, nucprior=2):
        """fastAlignmentRoutine with aln1/aln2/firstpos (reference: src/c/align.c:523-577).
        Returns (score, aln1, aln2, firstpos)."""
        L = len(seq2)
        if len(seq1) < L + 15 or len(gap_open) < L + 15 or len(qual2) != L:
            raise PlbError(_abi.PLB_ERR_SHAPE, "need len(seq1) >= len(seq2)+15 and matching qual/gap-open lengths")
        a1 = C.create_string_buffer(2 * L + 16)
        a2 = C.create_string_buffer(2 * L + 16)
        fp = C.c_int(0)
        s = _check(self.lib, self.lib.plb_fast_align(self.ctx, seq1, seq2, qual2, L + 15, L, gapextend, nucprior,
                                                     gap_open, C.addressof(a1), C.addressof(a2), C.addressof(fp)))
        return s, a1.value, a2.value, fp.value

    def align_traceback_batch(self, hap_segs, gap_opens, reads, quals, gapextend=3, nucprior=2):
        """Batched traceback: list of (score, aln1, aln2, firstpos)."""
        n = len(reads)
        ho = np.zeros(n + 1, np.int64)
        ro = np.zeros(n + 1, np.int64)
        np.cumsum([len(x) for x in hap_segs], out=ho[1:])
        np.cumsum([len(x) for x in reads], out=ro[1:])
        hseq = np.frombuffer(b"".join(hap_segs), np.uint8)
        gop = np.frombuffer(b"".join(g[:len(h)] for g, h in zip(gap_opens, hap_segs)), np.uint8)
        rseq = np.frombuffer(b"".join(reads), np.uint8)
        rq = np.frombuffer(b"".join(quals), np.uint8)
        sc = np.zeros(max(n, 1), np.int32)
        fp = np.zeros(max(n, 1), np.int32)
        nb = 2 * int(ro[-1]) + 16 * n
        a1 = np.zeros(max(nb, 1), np.uint8)
        a2 = np.zeros(max(nb, 1), np.uint8)
        _check(self.lib, self.lib.plb_align_traceback_host(self.ctx, n, _abi.ptr(ho), hseq.ctypes.data, gop.ctypes.data,
                                                           _abi.ptr(ro), rseq.ctypes.data, rq.ctypes.data, gapextend,
                                                           nucprior, _abi.ptr(sc), _abi.ptr(a1), _abi.ptr(a2),
                                                           _abi.ptr(fp)))
        out = []
        for i in range(n):
            o = 2 * int(ro[i]) + 16 * i
            r1 = a1[o:o + 2 * len(reads[i]) + 16].tobytes()
            r2 = a2[o:o + 2 * len(reads[i]) + 16].tobytes()
            out.append((int(sc[i]), r1[:r1.index(b"\0")], r2[:r2.index(b"\0")], int(fp[i])))
        return out

    def align_flank_batch(self, haps, gap_opens, seg_starts, hap_flanks, reads, quals, gapextend=3, nucprior=2):
        """fastAlignmentRoutine + calculateFlankScore (reference: src/c/align.c:593-644) for explicit
        (whole haplotype, band start, flank, read) tuples.  Returns (scores, flank scores)."""
        n = len(reads)
        ho = np.zeros(n + 1, np.int64)
        ro = np.zeros(n + 1, np.int64)
        np.cumsum([len(x) for x in haps], out=ho[1:])
        np.cumsum([len(x) for x in reads], out=ro[1:])
        hseq = np.frombuffer(b"".join(haps), np.uint8)
        gop = np.frombuffer(b"".join(g[:len(h)] for g, h in zip(gap_opens, haps)), np.uint8)
        rseq = np.frombuffer(b"".join(reads), np.uint8)
        rq = np.frombuffer(b"".join(quals), np.uint8)
        ss = np.asarray(seg_starts, np.int32)
        hf = np.asarray(hap_flanks, np.int32)
        sc = np.zeros(max(n, 1), np.int32)
        fl = np.zeros(max(n, 1), np.int32)
        _check(self.lib, self.lib.plb_align_flank_batch_host(self.ctx, n, _abi.ptr(ho), hseq.ctypes.data, gop.ctypes.data,
                                                             _abi.ptr(ss), _abi.ptr(hf), _abi.ptr(ro), rseq.ctypes.data,
                                                             rq.ctypes.data, gapextend, nucprior, _abi.ptr(sc),
                                                             _abi.ptr(fl)))
        return sc[:n], fl[:n]

    def align_batch(self, hap_segs, gap_opens, reads, quals, gapextend=3, nucprior=2):
        """Batched S1 over explicit (segment, read) pairs given as lists of bytes."""
        n = len(reads)
        ho = np.zeros(n + 1, np.int64)
        ro = np.zeros(n + 1, np.int64)
        np.cumsum([len(x) for x in hap_segs], out=ho[1:])
        np.cumsum([len(x) for x in reads], out=ro[1:])
        for hs, go, rd, q in zip(hap_segs, gap_opens, reads, quals):
            if len(go) < len(hs) or len(q) != len(rd):
                raise PlbError(_abi.PLB_ERR_SHAPE, "gap-open shorter than segment or qual/read length mismatch")
        hseq = np.frombuffer(b"".join(hap_segs), np.uint8)
        gop = np.frombuffer(b"".join(g[:len(h)] for g, h in zip(gap_opens, hap_segs)), np.uint8)
        rseq = np.frombuffer(b"".join(reads), np.uint8)
        rq = np.frombuffer(b"".join(quals), np.uint8)
        out = np.zeros(max(n, 1), np.int32)
        _check(self.lib, self.lib.plb_align_batch_host(self.ctx, n, _abi.ptr(ho), hseq.ctypes.data, gop.ctypes.data,
                                                       _abi.ptr(ro), rseq.ctypes.data, rq.ctypes.data, gapextend,
                                                       nucprior, _abi.ptr(out)))
        return out[:n]

    def gap_open(self, haps):
        """Haplotype.annotateWithGapOpen for a list of haplotype byte strings
        (reference: src/cython/chaplotype.pyx:552-590); returns a list of bytes of length hapLen+1."""
        n = len(haps)
        off = np.zeros(n + 1, np.int64)
        np.cumsum([len(h) for h in haps], out=off[1:])
        seq = np.frombuffer(b"".join(haps), np.uint8)
        out = np.zeros(int(off[-1]) + n + 1, np.uint8)
        _check(self.lib, self.lib.plb_gap_open_host(self.ctx, n, _abi.ptr(off), seq.ctypes.data if len(seq) else None,
                                                    _abi.ptr(out)))
        return [out[int(off[h]) + h:int(off[h + 1]) + h + 1].tobytes() for h in range(n)]

    # ---- S2 ---------------------------------------------------------------------------------
    def window_loglik(self, batch: WindowBatch, opt=None, want_score=True):
        """Per-read log-likelihoods of every haplotype of every window
        (Haplotype.alignReads, reference: src/cython/chaplotype.pyx:306-377).
        Returns (ll, score) flat arrays in the layout given by batch.ll_offsets()."""
        opt = opt or _abi.PlbOptions.default()
        off = batch.ll_offsets()
        n = int(off[-1])
        ll = np.zeros(max(n, 1), np.float64)
        sc = np.zeros(max(n, 1), np.int32) if want_score else None
        out = _abi.PlbLoglikOut(_abi.ptr(off), _abi.ptr(ll), _abi.ptr(sc))
        s = batch.as_struct()
        _check(self.lib, self.lib.plb_window_loglik_host(self.ctx, C.byref(s), C.byref(opt), C.byref(out)))
        return ll[:n], (sc[:n] if want_score else None)

    # ---- S3 ---------------------------------------------------------------------------------
    @staticmethod
    def alloc_population_out(batch: WindowBatch, max_haps=None):
        W, nI = batch.n_windows, batch.n_individuals
        Hm = max_haps or batch.max_haps()
        Gm = Hm * (Hm + 1) // 2
        V = max(batch.max_variants, 1)
        return {"max_haps": Hm, "gl": np.zeros((W, nI, Gm)), "gl_log_max": np.zeros((W, nI)),
                "gof": np.zeros((W, Gm, nI)), "hap_like": np.zeros((W, nI, Hm)), "freq": np.zeros((W, Hm)),
                "em_post": np.zeros((W, nI, Gm)), "call": np.zeros((W, nI), np.int32),
                "var_phred": np.zeros((W, V)), "em_iters": np.zeros(W, np.int32)}

    @staticmethod
    def _any_ptr(a):
        """Pointer of a numpy array (host) or of anything with data_ptr() (a torch tensor, host or device)."""
        if a is not None and not isinstance(a, np.ndarray) and hasattr(a, "data_ptr"):
            assert a.is_contiguous()
            return a.data_ptr()
        return _abi.ptr(a)

    @staticmethod
    def _pop_struct(arrs, ptr_of=None):
        ptr_of = ptr_of or Engine._any_ptr
        o = _abi.PlbPopulationOut()
        o.max_haps = arrs["max_haps"]
        for k in ("gl", "gl_log_max", "gof", "hap_like", "freq", "em_post", "call", "var_phred", "em_iters"):
            setattr(o, k, ptr_of(arrs.get(k)))
        return o

    def population_run(self, batch: WindowBatch, opt=None, max_haps=None, want_ll=False, out=None):
        """Population.setup + Population.call for every window (reference:
        src/cython/cpopulation.pyx:197-309, 678-720).  Host buffers in and out.
        Returns dict of arrays (+ 'll', 'score' when want_ll)."""
        opt = opt or _abi.PlbOptions.default()
        arrs = out or self.alloc_population_out(batch, max_haps)
        po = self._pop_struct(arrs)
        lo = None
        if want_ll:
            off = batch.ll_offsets()
            n = int(off[-1])
            ll = np.zeros(max(n, 1), np.float64)
            sc = np.zeros(max(n, 1), np.int32)
            lo = _abi.PlbLoglikOut(_abi.ptr(off), _abi.ptr(ll), _abi.ptr(sc))
        s = batch.as_struct()
        _check(self.lib, self.lib.plb_population_run_host(self.ctx, C.byref(s), C.byref(opt), C.byref(po),
                                                          C.byref(lo) if lo is not None else None))
        if want_ll:
            arrs["ll"], arrs["score"] = ll[:n], sc[:n]
        return arrs

    def population_submit(self, batch: WindowBatch, opt=None, max_haps=None, out=None):
        """First half of population_run for a region loop (variantcaller.pyx:566-615): queues the uploads, kernels and
        downloads of `batch` and returns a job handle without waiting.  Up to two jobs may be in flight per engine -
        batch k+1 is uploaded while batch k computes.  The arrays of `batch` and `out` must stay alive and untouched
        until population_wait(job); they should live in pinned memory (torch .pin_memory())."""
        opt = opt or _abi.PlbOptions.default()
        arrs = out or self.alloc_population_out(batch, max_haps)
        po = self._pop_struct(arrs)
        s = batch.as_struct()
        job = C.c_void_p()
        _check(self.lib, self.lib.plb_population_submit(self.ctx, C.byref(s), C.byref(opt), C.byref(po), None, C.byref(job)))
        return {"job": job, "out": arrs, "keep": (batch, s, po, opt)}

    def population_wait(self, handle):
        """Second half: blocks until the job's outputs are in its `out` arrays and returns them."""
        job, handle["job"] = handle["job"], None
        if job is not None:
            _check(self.lib, self.lib.plb_population_wait(self.ctx, job))
        handle["keep"] = None
        return handle["out"]

    # ---- N4 ---------------------------------------------------------------------------------
    def site_genotypes(self, batch: WindowBatch, pop: dict, sites):
        """computeGenotypeCallAndLikelihoods + the per-sample derivations of outputCallToVCF for every
        (site, individual) (reference: src/cython/vcfutils.pyx:163-334, 491-548).  `pop` is the dict
        population_run returned for `batch`; `sites` a batch.SiteBatch.  Returns a dict of arrays."""
        from .batch import alloc_site_out, site_out_struct
        arrs = alloc_site_out(batch, sites)
        so = site_out_struct(arrs)
        po = self._pop_struct(pop)
        s, ss = batch.as_struct(), sites.as_struct()
        _check(self.lib, self.lib.plb_site_genotypes_host(self.ctx, C.byref(s), C.byref(po), C.byref(ss), C.byref(so)))
        return arrs

    # ---- N1 ---------------------------------------------------------------------------------
    def build_haplotypes(self, ref_batch: WindowBatch, variants, hap_win, hap_mask):
        """Haplotype.haplotypeSequence of haplotype k = the reference haplotype of window hap_win[k] (ref_batch holds
        ONE haplotype per window) mutated by the window's variants in bit mask hap_mask[k] (reference:
        src/cython/chaplotype.pyx:127-172, 397-449).  Returns a list of bytes."""
        hap_win = np.ascontiguousarray(hap_win, np.int32)
        hap_mask = np.ascontiguousarray(hap_mask, np.uint64)
        n = len(hap_win)
        off = np.zeros(n + 1, np.int64)
        s, v = ref_batch.as_struct(), variants.as_struct()
        _check(self.lib, self.lib.plb_build_haplotypes_host(self.ctx, C.byref(s), C.byref(v), n, _abi.ptr(hap_win),
                                                            _abi.ptr(hap_mask), _abi.ptr(off), None, 0))
        seq = np.zeros(int(off[-1]) + 1, np.uint8)
        _check(self.lib, self.lib.plb_build_haplotypes_host(self.ctx, C.byref(s), C.byref(v), n, _abi.ptr(hap_win),
                                                            _abi.ptr(hap_mask), _abi.ptr(off), _abi.ptr(seq), int(off[-1])))
        return [seq[int(off[k]):int(off[k + 1])].tobytes() for k in range(n)]

    def select_haplotypes(self, ref_batch: WindowBatch, variants, sel=None, opt=None, max_sel=None, out=None):
        """getFilteredHaplotypes for every window of a batch (reference: src/cython/variantFilter.pyx:377-506,
        237-283).  ref_batch: ONE haplotype per window (the reference haplotype) + the good reads per individual;
        variants: batch.VariantSet.  Returns dict: n_sel [W], sel_mask [W][max_sel] (bit v = window variant v),
        sel_score [W][max_sel], n_scored [W]."""
        sel = sel or _abi.PlbSelectOptions.default()
        opt = opt or _abi.PlbOptions.default()
        W = ref_batch.n_windows
        if max_sel is None:
            nv = int(np.max(np.diff(variants.win_var_off))) if W else 0
            max_sel = max(1, sel.max_haplotypes - 1, sel.original_max_haplotypes - 1, min(2 ** min(nv, 20) - 1, 1 << 16))
        arrs = out if out is not None and out["max_sel"] == max_sel else {
            "max_sel": max_sel, "n_sel": np.zeros(W, np.int32), "sel_mask": np.zeros((W, max_sel), np.uint64),
            "sel_score": np.full((W, max_sel), np.nan), "n_scored": np.zeros(W, np.int32)}
        o = _abi.PlbSelectOut(max_sel, _abi.ptr(arrs["n_sel"]), _abi.ptr(arrs["sel_mask"]), _abi.ptr(arrs["sel_score"]),
                              _abi.ptr(arrs["n_scored"]))
        s, v = ref_batch.as_struct(), variants.as_struct()
        _check(self.lib, self.lib.plb_select_haplotypes_host(self.ctx, C.byref(s), C.byref(v), C.byref(sel), C.byref(opt),
                                                             C.byref(o)))
        return arrs

    def best_score_haplotypes(self, batch: WindowBatch, opt=None):
        """computeBestScoreForHaplotype for every haplotype of the batch (reference: src/cython/variantFilter.pyx:212-234):
        sum of alignSingleRead over an individual's good reads, best individual.  Returns [n_haps] float64."""
        opt = opt or _abi.PlbOptions.default()
        out = np.zeros(max(batch.n_haps, 1), np.float64)
        s = batch.as_struct()
        _check(self.lib, self.lib.plb_best_score_haplotypes_host(self.ctx, C.byref(s), C.byref(opt), _abi.ptr(out)))
        return out[:batch.n_haps]

    def best_score_genotypes(self, batch: WindowBatch, hap1, hap2, target_coverage=30, opt=None):
        """computeBestScoreForGenotype for pairs of haplotypes of the batch (reference: src/cython/variantFilter.pyx:237-283;
        the second pass of getAllHLAHaplotypesInRegion, :723-733).  hap1 / hap2: haplotype indices, pairwise in one window.
        Returns [n_pairs] float64."""
        opt = opt or _abi.PlbOptions.default()
        hap1 = np.ascontiguousarray(hap1, np.int32)
        hap2 = np.ascontiguousarray(hap2, np.int32)
        assert hap1.shape == hap2.shape and hap1.ndim == 1
        out = np.zeros(max(len(hap1), 1), np.float64)
        s = batch.as_struct()
        _check(self.lib, self.lib.plb_best_score_genotypes_host(self.ctx, C.byref(s), C.byref(opt), int(target_coverage),
                                                                len(hap1), _abi.ptr(hap1), _abi.ptr(hap2), _abi.ptr(out)))
        return out[:len(hap1)]

    def call_windows(self, ref_batch: WindowBatch, variants, sel=None, opt=None, var_prior=None):
        """callVariantsInWindow for a batch of windows (reference: src/cython/variantcaller.pyx:74-141): haplotype
        selection (getHaplotypesInWindow), haplotype construction, then Population.setup + call on
        [reference haplotype] + selected haplotypes.  Returns (selection dict, the window-model batch, population dict)."""
        from .batch import with_haplotypes
        sel_out = self.select_haplotypes(ref_batch, variants, sel, opt)
        W = ref_batch.n_windows
        n_sel = sel_out["n_sel"].astype(np.int64)
        hap_off = np.zeros(W + 1, np.int64)
        np.cumsum(n_sel + 1, out=hap_off[1:])
        hap_win = np.repeat(np.arange(W, dtype=np.int32), n_sel + 1)
        hap_mask = np.zeros(int(hap_off[-1]), np.uint64)          # mask 0 = the reference haplotype, first in every window
        for w in range(W):
            hap_mask[hap_off[w] + 1:hap_off[w + 1]] = sel_out["sel_mask"][w, :n_sel[w]]
        seqs = self.build_haplotypes(ref_batch, variants, hap_win, hap_mask)
        batch = with_haplotypes(ref_batch, hap_off, seqs, hap_mask, variants, var_prior)
        return sel_out, batch, self.population_run(batch, opt)

    @staticmethod
    def select_replay(ref_batch: WindowBatch, variants, score_fn, sel=None, max_sel=None, lib=None):
        """The selection loop's bookkeeping alone (plb_select_replay_host): score_fn(hap_win, hap_mask) -> scores is
        called once per round with numpy arrays.  No GPU work; needs no context."""
        lib = lib or load_library()
        sel = sel or _abi.PlbSelectOptions.default()
        W = ref_batch.n_windows
        if max_sel is None:
            nv = int(np.max(np.diff(variants.win_var_off))) if W else 0
            max_sel = max(1, sel.max_haplotypes - 1, sel.original_max_haplotypes - 1, min(2 ** min(nv, 20) - 1, 1 << 16))
        arrs = {"max_sel": max_sel, "n_sel": np.zeros(W, np.int32), "sel_mask": np.zeros((W, max_sel), np.uint64),
                "sel_score": np.full((W, max_sel), np.nan), "n_scored": np.zeros(W, np.int32)}
        o = _abi.PlbSelectOut(max_sel, _abi.ptr(arrs["n_sel"]), _abi.ptr(arrs["sel_mask"]), _abi.ptr(arrs["sel_score"]),
                              _abi.ptr(arrs["n_scored"]))

        def cb(user, n, hw, hm, out):
            try:
                sc = score_fn(np.ctypeslib.as_array(hw, (n,)).copy(), np.ctypeslib.as_array(hm, (n,)).copy())
                np.ctypeslib.as_array(out, (n,))[:] = sc
                return 0
            except Exception:   # surfaces as PLB_ERR_ARG
                import traceback
                traceback.print_exc()
                return 1
        s, v = ref_batch.as_struct(), variants.as_struct()
        _check(lib, lib.plb_select_replay_host(C.byref(s), C.byref(v), C.byref(sel), _abi.TRIAL_SCORE_FN(cb), None, C.byref(o)))
        return arrs

    def select_stats(self):
        """Stage times and counts of the last select_haplotypes (plb_select_stats)."""
        a = (C.c_double * 10)()
        _check(self.lib, self.lib.plb_select_stats(self.ctx, a, 10))
        keys = ("ref_pass_ms", "build_ms", "score_ms", "reduce_ms", "host_ms", "rounds", "n_trials", "n_pairs", "cells",
                "n_filter_windows")
        return dict(zip(keys, [float(x) for x in a]))

    # ---- device-resident path -------------------------------------------------------------------
    def upload(self, batch: WindowBatch):
        h = C.c_void_p()
        s = batch.as_struct()
        _check(self.lib, self.lib.plb_batch_upload(self.ctx, C.byref(s), C.byref(h)))
        return h

    def free(self, handle):
        self.lib.plb_batch_free(self.ctx, handle)

    def synth_fill(self, handle, first_window, seed=20261017):
        """Measurement support: refill a resident batch in place with device-generated synthetic windows
        (plb_synth_fill_device, "synth-v1d"); asynchronous on the context stream."""
        _check(self.lib, self.lib.plb_synth_fill_device(self.ctx, handle, int(seed), int(first_window)))

    def download(self, handle, batch: WindowBatch):
        """Copy the input arrays of a resident batch back into `batch` (same shape as the uploaded one), in place."""
        s = batch.as_struct()
        _check(self.lib, self.lib.plb_batch_download(self.ctx, handle, C.byref(s)))
        return batch

    def run_device(self, handle, pop_ptrs=None, ll_ptr=None, score_ptr=None, ll_off_ptr=None, opt=None):
        """Launch the whole path on a device-resident batch; asynchronous on the context stream.
        pop_ptrs: dict name -> device pointer (int) incl. 'max_haps'."""
        opt = opt or _abi.PlbOptions.default()
        po = None
        if pop_ptrs is not None:
            po = _abi.PlbPopulationOut()
            po.max_haps = pop_ptrs["max_haps"]
            for k in ("gl", "gl_log_max", "gof", "hap_like", "freq", "em_post", "call", "var_phred", "em_iters"):
                setattr(po, k, pop_ptrs.get(k))
        lo = _abi.PlbLoglikOut(ll_off_ptr, ll_ptr, score_ptr)
        _check(self.lib, self.lib.plb_run_device(self.ctx, handle, C.byref(opt), C.byref(po) if po is not None else None,
                                                 C.byref(lo)))
