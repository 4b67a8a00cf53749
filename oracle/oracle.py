"""
ctypes front-end of the CPU oracle (oracle/platypus_oracle.c) and of the compiled reference
(oracle/_ref/).  TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never from platypus_b200/.
"""
import ctypes as C
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from platypus_b200 import _abi  # noqa: E402  (struct layouts only)
from oracle import build as _build  # noqa: E402

_lib = None
_ref = None
_ALIGN_FN = C.CFUNCTYPE(C.c_int, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                        C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        path = _build.build_oracle()
        L = C.CDLL(path)
        u8 = C.c_char_p
        L.plo_band_align.argtypes = [u8, u8, u8, C.c_int, C.c_int, C.c_int, u8]
        L.plo_band_align.restype = C.c_int
        L.plo_gap_open.argtypes = [u8, C.c_int, C.c_void_p]
        L.plo_gap_open.restype = None
        L.plo_kmer_hash.argtypes = [u8]
        L.plo_kmer_hash.restype = C.c_uint32
        L.plo_map_and_align.argtypes = [u8, u8, C.c_int, C.c_int, C.c_int, C.c_int, u8, u8, C.c_int, C.c_int,
                                        C.POINTER(C.c_int)]
        L.plo_map_and_align.restype = C.c_int
        L.plo_score_to_ll.argtypes = [C.c_int, C.c_int]
        L.plo_score_to_ll.restype = C.c_double
        L.plo_score_to_ll_hla.argtypes = [C.c_int, C.c_int]
        L.plo_score_to_ll_hla.restype = C.c_double
        L.plo_band_align_tb.argtypes = [u8, u8, u8, C.c_int, C.c_int, C.c_int, u8, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_int)]
        L.plo_band_align_tb.restype = C.c_int
        L.plo_flank_score.argtypes = [C.c_int, C.c_int, u8, u8, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p]
        L.plo_flank_score.restype = C.c_int
        L.plo_band_align_flank.argtypes = [u8, u8, u8, C.c_int, C.c_int, C.c_int, u8, C.c_int, C.c_int, C.c_int,
                                           C.POINTER(C.c_int)]
        L.plo_band_align_flank.restype = C.c_int
        L.plo_map_and_align_ex.argtypes = [u8, u8, C.c_int, C.c_int, C.c_int, C.c_int, u8, u8, C.c_int, C.c_int,
                                           C.c_int, C.c_int, u8, C.c_int, C.POINTER(C.c_int)]
        L.plo_map_and_align_ex.restype = C.c_int
        L.plo_set_flank_fn.argtypes = [C.c_void_p]
        L.plo_set_flank_fn.restype = None
        L.plo_overlap.argtypes = [C.c_int] * 4
        L.plo_overlap.restype = C.c_int
        L.plo_window_loglik.argtypes = [C.POINTER(_abi.PlbWindowBatch), C.POINTER(_abi.PlbOptions),
                                        C.POINTER(_abi.PlbLoglikOut), C.c_int, C.POINTER(_abi.PlbRunStats)]
        L.plo_window_loglik.restype = C.c_int
        L.plo_population_run.argtypes = [C.POINTER(_abi.PlbWindowBatch), C.POINTER(_abi.PlbOptions),
                                         C.POINTER(_abi.PlbPopulationOut), C.POINTER(_abi.PlbLoglikOut), C.c_int,
                                         C.POINTER(_abi.PlbRunStats)]
        L.plo_population_run.restype = C.c_int
        L.plo_genotype_loglik.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double),
                                          C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.plo_genotype_loglik.restype = C.c_double
        L.plo_em.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.plo_em.restype = C.c_int
        L.plo_posterior.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_double]
        L.plo_posterior.restype = C.c_double
        L.plo_site_genotypes.argtypes = [C.POINTER(_abi.PlbWindowBatch), C.POINTER(_abi.PlbPopulationOut),
                                         C.POINTER(_abi.PlbSiteBatch), C.POINTER(_abi.PlbSiteOut)]
        L.plo_site_genotypes.restype = C.c_int
        L.plo_set_align_fn.argtypes = [C.c_void_p, C.c_int]
        L.plo_set_align_fn.restype = None
        _lib = L
    return _lib


# ---- the compiled reference (L1) -----------------------------------------------------------
def ref_align_lib():
    """oracle/_ref/libalign_ref.so = unmodified src/c/align.c, or None if never built."""
    global _ref
    if _ref is None:
        path = _build.build_align_ref()
        if path is None or not os.path.exists(path):
            return None
        R = C.CDLL(path)
        R.fastAlignmentRoutine.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_char_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int)]
        R.fastAlignmentRoutine.restype = C.c_int
        R.calculateFlankScore.argtypes = [C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int,
                                          C.c_char_p, C.c_char_p]
        R.calculateFlankScore.restype = C.c_int
        _ref = R
    return _ref


def ref_fast_align(hap_seg: bytes, read: bytes, qual: bytes, gap_open: bytes, ext=3, nuc=2, traceback=False):
    """fastAlignmentRoutine of the reference (src/c/align.c:77)."""
    R = ref_align_lib()
    L = len(read)
    assert len(hap_seg) >= L + 15 and len(gap_open) >= L + 15
    fp = C.c_int(0)
    if traceback:
        a1 = C.create_string_buffer(2 * L + 16)
        a2 = C.create_string_buffer(2 * L + 16)
        return R.fastAlignmentRoutine(hap_seg, read, qual, L + 15, L, ext, nuc, gap_open, a1, a2, C.byref(fp))
    return R.fastAlignmentRoutine(hap_seg, read, qual, L + 15, L, ext, nuc, gap_open, None, None, C.byref(fp))


def ref_fast_align_tb(hap_seg: bytes, read: bytes, qual: bytes, gap_open: bytes, ext=3, nuc=2):
    """fastAlignmentRoutine of the reference with traceback: (score, aln1, aln2, firstpos)."""
    R = ref_align_lib()
    L = len(read)
    assert len(hap_seg) >= L + 15 and len(gap_open) >= L + 15
    fp = C.c_int(0)
    a1 = C.create_string_buffer(2 * L + 16)
    a2 = C.create_string_buffer(2 * L + 16)
    s = R.fastAlignmentRoutine(hap_seg, read, qual, L + 15, L, ext, nuc, gap_open, a1, a2, C.byref(fp))
    return s, a1.value, a2.value, fp.value


def ref_flank_score(hap_len, hap_flank, qual: bytes, gap_open: bytes, firstpos, aln1: bytes, aln2: bytes, ext=3, nuc=2):
    """calculateFlankScore of the reference (src/c/align.c:593)."""
    return ref_align_lib().calculateFlankScore(hap_len, hap_flank, qual, gap_open, ext, nuc, firstpos, aln1, aln2)


def band_align_tb(hap_seg: bytes, read: bytes, qual: bytes, gap_open: bytes, ext=3, nuc=2):
    """Restated traceback: (score, aln1, aln2, firstpos)."""
    L = len(read)
    assert len(hap_seg) >= L + 15 and len(gap_open) >= L + 15
    fp = C.c_int(0)
    a1 = C.create_string_buffer(2 * L + 16)
    a2 = C.create_string_buffer(2 * L + 16)
    s = lib().plo_band_align_tb(hap_seg, read, qual, L, ext, nuc, gap_open, a1, a2, C.byref(fp))
    return s, a1.value, a2.value, fp.value


def flank_score(hap_len, hap_flank, qual: bytes, gap_open: bytes, firstpos, aln1: bytes, aln2: bytes, ext=3, nuc=2):
    return lib().plo_flank_score(hap_len, hap_flank, qual, gap_open, ext, nuc, firstpos, aln1, aln2)


def band_align_flank(hap: bytes, read: bytes, qual: bytes, gap_open: bytes, start, hap_flank, ext=3, nuc=2):
    """One forward pass: (score, flank score of the traceback alignment).  hap / gap_open are the
    WHOLE haplotype and its table; the band segment starts at `start`."""
    L = len(read)
    assert len(hap) >= start + L + 15
    f = C.c_int(0)
    s = lib().plo_band_align_flank(hap[start:], read, qual, L, ext, nuc, gap_open[start:], start, len(hap),
                                   hap_flank, C.byref(f))
    return s, f.value


def use_reference_kernel(on=True, traceback=True):
    """Route every band alignment of the oracle through the reference's align.c
    (kind 'reference' CPU baseline).  Returns False when oracle/_ref is absent."""
    L = lib()
    if not on:
        L.plo_set_align_fn(None, 0)
        L.plo_set_flank_fn(None)
        return True
    R = ref_align_lib()
    if R is None:
        return False
    L.plo_set_align_fn(C.cast(R.fastAlignmentRoutine, C.c_void_p), 1 if traceback else 0)
    L.plo_set_flank_fn(C.cast(R.calculateFlankScore, C.c_void_p))
    return True


def ref_calign():
    """The reference's calign.pyx behind oracle/calign_ref_wrap.pyx, or None."""
    try:
        paths = _build.build_calign_ref()
    except Exception:
        paths = None
    if not paths:
        return None
    d = os.path.dirname(paths[0])
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return importlib.import_module("calign_ref_wrap")
    except ImportError:
        return None


def ref_l3():
    """The reference's chaplotype.pyx / cgenotype.pyx behind oracle/l3_ref_wrap.pyx, or None."""
    try:
        paths = _build.build_l3_ref()
    except Exception:
        paths = None
    if not paths:
        return None
    d = os.path.dirname(paths[0])
    if d not in sys.path:
        sys.path.insert(0, d)
    try:
        return importlib.import_module("l3_ref_wrap")
    except ImportError:
        return None


def ref_n4():
    """computeGenotypeCallAndLikelihoods of the reference (vcfutils.pyx:163-334), excerpted at build time, or None."""
    if ref_l3() is None:
        return None
    try:
        return importlib.import_module("n4_ref")
    except ImportError:
        return None


def ref_site_genotypes(batch, pop_arrs, sites):
    """The reference's computeGenotypeCallAndLikelihoods for every (site, individual with reads) of a site batch:
    dict (site, individual) -> (phasedIndex1, phasedIndex2, likelihoods, genotype / non-ref / ref posterior, GOF)."""
    N = ref_n4()
    nI = batch.n_individuals
    Hm = pop_arrs["max_haps"]
    out = {}
    for s in range(sites.n_sites):
        w = int(sites.site_win[s])
        h0, h1 = int(batch.win_hap_off[w]), int(batch.win_hap_off[w + 1])
        H = h1 - h0
        G = H * (H + 1) // 2
        vs = [int(v) for v in sites.site_var[sites.site_var_off[s]:sites.site_var_off[s + 1]]]
        isref = [int(x) for x in sites.hap_is_ref[sites.site_hap_off[s]:sites.site_hap_off[s + 1]]]
        vih = [[int((int(batch.hap_var_mask[h0 + h]) >> v) & 1) for v in vs] for h in range(H)]
        for i in range(nI):
            if batch.wi_n_good[w * nI + i] == 0:
                continue
            out[(s, i)] = N.compute_genotype_call_and_likelihoods(
                len(vs), H, [float(x) for x in pop_arrs["freq"][w, :H]], [float(x) for x in pop_arrs["gl"][w, i, :G]],
                [float(x) for x in pop_arrs["gof"][w, :G, i]], vih, isref, nI)
    return out


# ---- restatement entry points -----------------------------------------------------------------
def band_align(hap_seg: bytes, read: bytes, qual: bytes, gap_open: bytes, ext=3, nuc=2):
    assert len(hap_seg) >= len(read) + 15 and len(gap_open) >= len(read) + 15
    return lib().plo_band_align(hap_seg, read, qual, len(read), ext, nuc, gap_open)


def gap_open(hap: bytes) -> bytes:
    out = C.create_string_buffer(len(hap) + 1)
    lib().plo_gap_open(hap, len(hap), out)
    return out.raw[:len(hap) + 1]


def kmer_hash(seq: bytes) -> int:
    return lib().plo_kmer_hash(seq)


def map_and_align(read: bytes, qual: bytes, read_start: int, hap_start: int, hap: bytes, go: bytes = None,
                  ext=3, nuc=2):
    if go is None:
        go = gap_open(hap)
    n = C.c_int(0)
    s = lib().plo_map_and_align(read, qual, read_start, hap_start, len(read), len(hap), hap, go, ext, nuc,
                                C.byref(n))
    return s, n.value


def map_and_align_ex(read: bytes, qual: bytes, read_start: int, hap_start: int, hap: bytes, go: bytes = None,
                     hap_flank=1, do_flank=0, hash_read: bytes = None, ext=3, nuc=2):
    if go is None:
        go = gap_open(hap)
    n = C.c_int(0)
    s = lib().plo_map_and_align_ex(read, qual, read_start, hap_start, len(read), len(hap), hap, go, ext, nuc,
                                   hap_flank, do_flank, hash_read, len(hash_read) if hash_read else 0, C.byref(n))
    return s, n.value


def score_to_ll(score, mapq):
    return lib().plo_score_to_ll(score, mapq)


def score_to_ll_hla(score, mapq):
    return lib().plo_score_to_ll_hla(score, mapq)


def window_loglik(batch, opt=None, n_threads=1, want_score=True):
    """Returns (ll, score, stats) with the PlbLoglikOut layout."""
    opt = opt or _abi.PlbOptions.default()
    off = batch.ll_offsets()
    n = int(off[-1])
    ll = np.zeros(max(n, 1), np.float64)
    sc = np.zeros(max(n, 1), np.int32) if want_score else None
    out = _abi.PlbLoglikOut(_abi.ptr(off), _abi.ptr(ll), _abi.ptr(sc))
    st = _abi.PlbRunStats()
    s = batch.as_struct()
    rc = lib().plo_window_loglik(C.byref(s), C.byref(opt), C.byref(out), n_threads, C.byref(st))
    if rc:
        raise RuntimeError("oracle plo_window_loglik failed: %d" % rc)
    return ll[:n], (sc[:n] if sc is not None else None), st.as_dict()


def alloc_population_out(batch, max_haps=None):
    W, nI = batch.n_windows, batch.n_individuals
    Hm = max_haps or batch.max_haps()
    Gm = Hm * (Hm + 1) // 2
    V = max(batch.max_variants, 1)
    return {
        "max_haps": Hm,
        "gl": np.zeros((W, nI, Gm)), "gl_log_max": np.zeros((W, nI)), "gof": np.zeros((W, Gm, nI)),
        "hap_like": np.zeros((W, nI, Hm)), "freq": np.zeros((W, Hm)), "em_post": np.zeros((W, nI, Gm)),
        "call": np.zeros((W, nI), np.int32), "var_phred": np.zeros((W, V)), "em_iters": np.zeros(W, np.int32),
    }


def population_struct(arrs):
    o = _abi.PlbPopulationOut()
    o.max_haps = arrs["max_haps"]
    for k in ("gl", "gl_log_max", "gof", "hap_like", "freq", "em_post", "call", "var_phred", "em_iters"):
        setattr(o, k, _abi.ptr(arrs[k]))
    return o


def population_run(batch, opt=None, n_threads=1, max_haps=None, want_ll=True):
    """Returns (dict of population outputs, ll, score, stats)."""
    opt = opt or _abi.PlbOptions.default()
    arrs = alloc_population_out(batch, max_haps)
    po = population_struct(arrs)
    off = batch.ll_offsets()
    n = int(off[-1])
    ll = np.zeros(max(n, 1), np.float64)
    sc = np.zeros(max(n, 1), np.int32)
    lo = _abi.PlbLoglikOut(_abi.ptr(off), _abi.ptr(ll), _abi.ptr(sc))
    st = _abi.PlbRunStats()
    s = batch.as_struct()
    rc = lib().plo_population_run(C.byref(s), C.byref(opt), C.byref(po), C.byref(lo) if want_ll else None,
                                  n_threads, C.byref(st))
    if rc:
        raise RuntimeError("oracle plo_population_run failed: %d" % rc)
    return arrs, ll[:n], sc[:n], st.as_dict()


def site_genotypes(batch, pop_arrs, sites):
    """N4 restatement: dict of per-(site, individual) arrays (layout of PlbSiteOut)."""
    from platypus_b200.batch import alloc_site_out, site_out_struct
    arrs = alloc_site_out(batch, sites)
    so = site_out_struct(arrs)
    po = population_struct(pop_arrs)
    s, ss = batch.as_struct(), sites.as_struct()
    rc = lib().plo_site_genotypes(C.byref(s), C.byref(po), C.byref(ss), C.byref(so))
    if rc:
        raise RuntimeError("oracle plo_site_genotypes failed: %d" % rc)
    return arrs
