"""ctypes mirror of include/platypus_b200.h (structs only; no library is loaded here)."""
import ctypes as C

import numpy as np

PLB_OK = 0
PLB_ERR_ARG = -1
PLB_ERR_SHAPE = -2
PLB_ERR_CUDA = -3
PLB_ERR_UNSUPPORTED = -4
PLB_ERR_NOMEM = -5
PLB_SCORE_NONE = 1000000
PLB_LL_CAP = -300.0
PLB_MAX_HAP_LEN = 16384
PLB_SEQ_ASCII = 0
PLB_SEQ_2BIT = 1
PLB_MAX_JOBS = 2

_p = C.c_void_p


class PlbOptions(C.Structure):
    _fields_ = [("gap_extend", C.c_int32), ("nuc_prior", C.c_int32), ("use_mapq_cap", C.c_int32),
                ("calc_flank_score", C.c_int32), ("max_em_iters", C.c_int32), ("use_em_likelihoods", C.c_int32)]

    @classmethod
    def default(cls, **kw):
        o = cls(3, 2, 0, 0, 100, 0)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class PlbWindowBatch(C.Structure):
    _fields_ = [("n_windows", C.c_int32), ("n_individuals", C.c_int32), ("n_haps", C.c_int32),
                ("n_reads", C.c_int32), ("n_slots", C.c_int64),
                ("win_hap_off", _p), ("win_start", _p), ("win_end", _p), ("hap_start", _p),
                ("hap_seq_off", _p), ("hap_seq", _p),
                ("wi_slot_off", _p), ("wi_n_good", _p), ("wi_n_bad", _p),
                ("slot_read", _p),
                ("read_seq_off", _p), ("read_seq", _p), ("read_qual", _p), ("read_pos", _p), ("read_end", _p),
                ("read_mapq", _p), ("read_qcfail", _p),
                ("max_variants", C.c_int32), ("win_n_var", _p), ("hap_var_mask", _p), ("var_prior", _p),
                ("seq_format", C.c_int32), ("n_read_exc", C.c_int64), ("read_exc_pos", _p), ("read_exc_chr", _p),
                ("n_hap_exc", C.c_int64), ("hap_exc_pos", _p), ("hap_exc_chr", _p),
                ("qual_bits", C.c_int32), ("qual_table", C.c_uint8 * 64)]


class PlbLoglikOut(C.Structure):
    _fields_ = [("ll_off", _p), ("ll", _p), ("score", _p)]


class PlbPopulationOut(C.Structure):
    _fields_ = [("max_haps", C.c_int32), ("gl", _p), ("gl_log_max", _p), ("gof", _p), ("hap_like", _p),
                ("freq", _p), ("em_post", _p), ("call", _p), ("var_phred", _p), ("em_iters", _p)]


class PlbSiteBatch(C.Structure):
    _fields_ = [("n_sites", C.c_int32), ("site_win", _p), ("site_var_off", _p), ("site_var", _p),
                ("site_hap_off", _p), ("hap_is_ref", _p), ("min_posterior", C.c_int32)]


class PlbSiteOut(C.Structure):
    _fields_ = [("max_pairs", C.c_int32), ("phased", _p), ("lik", _p), ("post", _p), ("phred", _p), ("gof", _p),
                ("gt", _p), ("gl_log10", _p)]


class PlbVariantSet(C.Structure):
    _fields_ = [("win_var_off", _p), ("var_pos", _p), ("var_n_removed", _p), ("var_n_support", _p),
                ("var_added_off", _p), ("var_added", _p)]


class PlbSelectOptions(C.Structure):
    _fields_ = [("max_haplotypes", C.c_int32), ("original_max_haplotypes", C.c_int32), ("max_variants", C.c_int32),
                ("filter_vars_by_coverage", C.c_int32), ("coverage_sampling_level", C.c_int32)]

    @classmethod
    def default(cls, **kw):
        """Defaults of src/python/runner.py:519-597 (originalMaxHaplotypes = maxHaplotypes, variantcaller.pyx:920)."""
        o = cls(50, 50, 8, 1, 30)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class PlbSelectOut(C.Structure):
    _fields_ = [("max_sel", C.c_int32), ("n_sel", _p), ("sel_mask", _p), ("sel_score", _p), ("n_scored", _p)]


class PlbBamRecords(C.Structure):
    _fields_ = [("n", C.c_int32), ("ref_id", _p), ("pos", _p), ("mapq", _p), ("flag", _p), ("mate_ref_id", _p),
                ("mate_pos", _p), ("tlen", _p), ("cigar_off", _p), ("cigar", _p), ("seq_off", _p), ("nib_off", _p),
                ("nib", _p), ("qual", _p)]


class PlbReadFilterOptions(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("min_good_qual_bases", "min_map_qual", "min_base_qual", "trim_read_flank",
                                         "trim_overlapping", "trim_adapter", "trim_soft_clipped", "filter_duplicates",
                                         "filter_mate_unmapped", "filter_mate_distant", "filter_small_insert")]

    @classmethod
    def default(cls, **kw):
        """Defaults of src/python/runner.py:551-580."""
        o = cls(20, 20, 20, 0, 1, 1, 1, 1, 1, 1, 1)
        for k, v in kw.items():
            setattr(o, k, v)
        return o


class PlbStagedReads(C.Structure):
    _fields_ = [("kept", _p), ("good", _p), ("read_pos", _p), ("read_end", _p), ("flag_out", _p), ("qual_out", _p),
                ("seq2", _p), ("exc_cap", C.c_int64), ("n_exc", C.c_int64), ("exc_pos", _p), ("exc_chr", _p),
                ("counts", C.c_int32 * 7)]


TRIAL_SCORE_FN = C.CFUNCTYPE(C.c_int, _p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_double))


class PlbRunStats(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("n_pairs_scored", C.c_int64), ("n_dp", C.c_int64), ("cells", C.c_int64),
                ("n_anchor_heavy", C.c_int64), ("n_anchor_verify", C.c_int64), ("n_anchor_exact", C.c_int64)]

    def as_dict(self):
        return {k: int(getattr(self, k)) for k, _ in self._fields_}


def ptr(a):
    """Host pointer of a C-contiguous numpy array (None -> NULL)."""
    if a is None:
        return None
    assert isinstance(a, np.ndarray) and a.flags["C_CONTIGUOUS"], "need a C-contiguous numpy array"
    return a.ctypes.data


def declare(lib):
    """Attach argtypes/restypes for every symbol of include/platypus_b200.h."""
    P = C.POINTER
    lib.plb_context_create.argtypes = [C.c_int, _p, P(_p)]
    lib.plb_context_create.restype = C.c_int
    lib.plb_context_destroy.argtypes = [_p]
    lib.plb_context_destroy.restype = None
    lib.plb_last_error.argtypes = []
    lib.plb_last_error.restype = C.c_char_p
    lib.plb_abi_version.argtypes = []
    lib.plb_abi_version.restype = C.c_int
    lib.plb_launch_count.argtypes = [_p]
    lib.plb_launch_count.restype = C.c_int64
    lib.plb_ll_offsets.argtypes = [P(PlbWindowBatch), _p, P(C.c_int64)]
    lib.plb_ll_offsets.restype = C.c_int
    lib.plb_validate.argtypes = [P(PlbWindowBatch), P(PlbOptions), C.c_int32]
    lib.plb_validate.restype = C.c_int
    lib.plb_fast_align.argtypes = [_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_char_p, _p, _p, _p]
    lib.plb_fast_align.restype = C.c_int
    lib.plb_align_batch_host.argtypes = [_p, C.c_int32, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, _p]
    lib.plb_align_batch_host.restype = C.c_int
    lib.plb_align_traceback_host.argtypes = [_p, C.c_int32, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p, _p, _p]
    lib.plb_align_traceback_host.restype = C.c_int
    lib.plb_align_flank_batch_host.argtypes = [_p, C.c_int32, _p, _p, _p, _p, _p, _p, _p, _p, C.c_int, C.c_int, _p, _p]
    lib.plb_align_flank_batch_host.restype = C.c_int
    lib.plb_gap_open_host.argtypes = [_p, C.c_int32, _p, _p, _p]
    lib.plb_gap_open_host.restype = C.c_int
    lib.plb_window_loglik_host.argtypes = [_p, P(PlbWindowBatch), P(PlbOptions), P(PlbLoglikOut)]
    lib.plb_window_loglik_host.restype = C.c_int
    lib.plb_population_run_host.argtypes = [_p, P(PlbWindowBatch), P(PlbOptions), P(PlbPopulationOut), P(PlbLoglikOut)]
    lib.plb_population_run_host.restype = C.c_int
    lib.plb_population_submit.argtypes = [_p, P(PlbWindowBatch), P(PlbOptions), P(PlbPopulationOut), P(PlbLoglikOut), P(_p)]
    lib.plb_population_submit.restype = C.c_int
    lib.plb_population_wait.argtypes = [_p, _p]
    lib.plb_population_wait.restype = C.c_int
    lib.plb_pack_bases_host.argtypes = [_p, C.c_int64, _p, C.c_int64, _p, _p, C.c_int64, P(C.c_int64)]
    lib.plb_pack_bases_host.restype = C.c_int
    lib.plb_pack_nibbles_host.argtypes = [_p, C.c_int64, _p, C.c_int64, _p, _p, C.c_int64, P(C.c_int64)]
    lib.plb_pack_nibbles_host.restype = C.c_int
    lib.plb_pack_quals_host.argtypes = [_p, C.c_int64, _p, P(C.c_int32), _p]
    lib.plb_pack_quals_host.restype = C.c_int
    lib.plb_stage_reads_host.argtypes = [P(PlbBamRecords), P(PlbReadFilterOptions), P(PlbStagedReads)]
    lib.plb_stage_reads_host.restype = C.c_int
    lib.plb_window_slices_host.argtypes = [C.c_int32, _p, _p, C.c_int32, _p, _p, _p, _p]
    lib.plb_window_slices_host.restype = C.c_int
    lib.plb_site_genotypes_host.argtypes = [_p, P(PlbWindowBatch), P(PlbPopulationOut), P(PlbSiteBatch), P(PlbSiteOut)]
    lib.plb_site_genotypes_host.restype = C.c_int
    lib.plb_batch_upload.argtypes = [_p, P(PlbWindowBatch), P(_p)]
    lib.plb_batch_upload.restype = C.c_int
    lib.plb_batch_free.argtypes = [_p, _p]
    lib.plb_batch_free.restype = None
    lib.plb_batch_download.argtypes = [_p, _p, P(PlbWindowBatch)]
    lib.plb_batch_download.restype = C.c_int
    lib.plb_synth_fill_device.argtypes = [_p, _p, C.c_uint64, C.c_int64]
    lib.plb_synth_fill_device.restype = C.c_int
    lib.plb_run_device.argtypes = [_p, _p, P(PlbOptions), P(PlbPopulationOut), P(PlbLoglikOut)]
    lib.plb_run_device.restype = C.c_int
    lib.plb_last_stats.argtypes = [_p, P(PlbRunStats)]
    lib.plb_last_stats.restype = C.c_int
    lib.plb_set_timing.argtypes = [_p, C.c_int]
    lib.plb_set_timing.restype = C.c_int
    lib.plb_kernel_times.argtypes = [_p, P(C.c_float)]
    lib.plb_kernel_times.restype = C.c_int
    lib.plb_build_haplotypes_host.argtypes = [_p, P(PlbWindowBatch), P(PlbVariantSet), C.c_int32, _p, _p, _p, _p, C.c_int64]
    lib.plb_build_haplotypes_host.restype = C.c_int
    lib.plb_select_haplotypes_host.argtypes = [_p, P(PlbWindowBatch), P(PlbVariantSet), P(PlbSelectOptions), P(PlbOptions),
                                               P(PlbSelectOut)]
    lib.plb_select_haplotypes_host.restype = C.c_int
    lib.plb_best_score_haplotypes_host.argtypes = [_p, P(PlbWindowBatch), P(PlbOptions), _p]
    lib.plb_best_score_haplotypes_host.restype = C.c_int
    lib.plb_best_score_genotypes_host.argtypes = [_p, P(PlbWindowBatch), P(PlbOptions), C.c_int32, C.c_int32, _p, _p, _p]
    lib.plb_best_score_genotypes_host.restype = C.c_int
    lib.plb_select_replay_host.argtypes = [P(PlbWindowBatch), P(PlbVariantSet), P(PlbSelectOptions), TRIAL_SCORE_FN, _p,
                                           P(PlbSelectOut)]
    lib.plb_select_replay_host.restype = C.c_int
    lib.plb_select_stats.argtypes = [_p, P(C.c_double), C.c_int]
    lib.plb_select_stats.restype = C.c_int
    return lib


# every symbol the header declares; tests check the built library exports all of them
EXPORTED_SYMBOLS = [
    "plb_context_create", "plb_context_destroy", "plb_last_error", "plb_abi_version", "plb_launch_count",
    "plb_ll_offsets", "plb_validate", "plb_fast_align", "plb_align_batch_host", "plb_align_traceback_host",
    "plb_align_flank_batch_host", "plb_gap_open_host",
    "plb_window_loglik_host", "plb_population_run_host", "plb_site_genotypes_host", "plb_batch_upload", "plb_batch_free",
    "plb_run_device", "plb_last_stats", "plb_set_timing", "plb_kernel_times",
    "plb_build_haplotypes_host", "plb_select_haplotypes_host", "plb_select_replay_host", "plb_best_score_haplotypes_host", "plb_best_score_genotypes_host",
    "plb_select_stats", "plb_population_submit", "plb_population_wait", "plb_pack_bases_host", "plb_pack_nibbles_host", "plb_pack_quals_host",
    "plb_stage_reads_host", "plb_window_slices_host", "plb_batch_download", "plb_synth_fill_device",
]
KERNEL_NAMES = ["k_prep", "k_anchor", "k_general", "k_dp", "k_genotype", "k_population"]
