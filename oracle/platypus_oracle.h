/*
 * platypus_oracle.h — CPU restatement of the Platypus read-vs-haplotype likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in platypus_b200/ may include, link or call this.
 * It exists so tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs can check (and time) the CUDA path against the reference algorithm.
 *
 * Parity pinning: the reference ships no golden vectors for this path.
 * The restatement is pinned by executing the reference itself in the build container:
 *   L1  plo_band_align, plo_band_align_tb, plo_flank_score   vs unmodified src/c/align.c
 *                                                             (oracle/_ref/libalign_ref.so)
 *   L2  plo_map_and_align(_ex)                                vs src/cython/calign.pyx
 *                                                             (oracle/_ref/calign_ref_wrap*.so)
 *   L3  per-read log-likelihoods (Haplotype.alignReads, default / HLA / flank mode), genotype
 *       log-likelihoods, GOF and hapLike (DiploidGenotype.calculateDataLikelihood), and the window model
 *       (Population.setup + call: max-rescale, EM frequencies and posteriors, genotype calls,
 *       calculatePosterior)                                   vs src/cython/chaplotype.pyx, cgenotype.pyx,
 *                                                             cpopulation.pyx (oracle/_ref/l3_ref_wrap*.so)
 * and the resulting inputs/outputs are committed as tests/golden/ *.npz.
 *   N4  plo_site_genotypes: phased indices, marginal likelihoods, posteriors, best GOF
 *                                                             vs computeGenotypeCallAndLikelihoods, the function's
 *                                                             own lines of src/cython/vcfutils.pyx:163-334 excerpted
 *                                                             at build time (oracle/_ref/n4_ref*.so)
 *   N1  oracle/select_oracle.py (Python: haplotype construction, isHaplotypeValid, computeBestScoreForGenotype,
 *       getFilteredHaplotypes)                                vs the functions' own lines of src/cython/variantFilter.pyx:237-283,
 *                                                             377-506 and platypusutils.pyx:735-802 excerpted at build time
 *                                                             (oracle/_ref/n1_ref*.so), on the reference's Variant / Haplotype objects
 * Still restated without a reference run ("parity unpinned"): only the few lines of outputCallToVCF that turn
 * those posteriors into phred values, the GT fallback rules and log10 GLs (vcfutils.pyx:504-548; inline in a
 * function that writes VCF through the reference's Python-2 I/O stack).
 */
#ifndef PLATYPUS_ORACLE_H
#define PLATYPUS_ORACLE_H

#include <stdint.h>
#include "../include/platypus_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* signature of the reference kernel, src/c/align.h:8-9 */
typedef int (*plo_align_fn)(const char* seq1, const char* seq2, const char* qual2, int len1, int len2,
                            int gapextend, int nucprior, const char* localgapopen,
                            char* aln1, char* aln2, int* firstpos);

/* Use `fn` (e.g. fastAlignmentRoutine dlsym'd from oracle/_ref/libalign_ref.so) for every
 * band alignment instead of plo_band_align; NULL restores the restatement.
 * `traceback` != 0 passes aln buffers so the reference also does its traceback, as
 * production does (calign.pyx:199-202). */
void plo_set_align_fn(plo_align_fn fn, int traceback);

/* L1: src/c/align.c:77-586 restated in matrix coordinates. */
int plo_band_align(const uint8_t* hap_seg, const uint8_t* read, const uint8_t* qual, int read_len,
                   int gap_extend, int nuc_prior, const uint8_t* gap_open);

/* L1 with traceback (src/c/align.c:344-365, 493-577): aln1/aln2 receive the NUL-terminated
 * alignment rows (2*read_len+16 bytes each), *firstpos the segment column of the first row. */
int plo_band_align_tb(const uint8_t* hap_seg, const uint8_t* read, const uint8_t* qual, int read_len,
                      int gap_extend, int nuc_prior, const uint8_t* gap_open, char* aln1, char* aln2,
                      int* firstpos);

/* src/c/align.c:593-644 calculateFlankScore. */
int plo_flank_score(int hap_len, int hap_flank, const uint8_t* qual, const uint8_t* gap_open, int gap_extend,
                    int nuc_prior, int firstpos, const char* aln1, const char* aln2);
typedef int (*plo_flank_fn)(int hapLen, int hapFlank, const char* quals, const char* localgapopen, int gapextend,
                            int nucprior, int firstpos, const char* aln1, const char* aln2);
/* Use the reference's own calculateFlankScore (oracle/_ref/libalign_ref.so); NULL restores ours. */
void plo_set_flank_fn(plo_flank_fn fn);

/* Traceback + flank score in ONE forward pass (the formulation the CUDA path uses): returns the
 * score, *flank = calculateFlankScore of the alignment the traceback would have produced.
 * `start` = offset of hap_seg inside the haplotype (gap_open is the segment's table). */
int plo_band_align_flank(const uint8_t* hap_seg, const uint8_t* read, const uint8_t* qual, int read_len,
                         int gap_extend, int nuc_prior, const uint8_t* gap_open, int start, int hap_len,
                         int hap_flank, int* flank);

/* src/cython/chaplotype.pyx:64-67, 552-590: out[hap_len+1]. */
void plo_gap_open(const uint8_t* hap, int hap_len, uint8_t* out);

/* src/cython/calign.pyx:61-90 */
uint32_t plo_kmer_hash(const uint8_t* seq);

/* L2: src/cython/calign.pyx:170-272.  n_dp (may be NULL) receives the number of band
 * alignments the reference would have executed for this pair. */
int plo_map_and_align(const uint8_t* read, const uint8_t* qual, int read_start, int hap_start,
                      int read_len, int hap_len, const uint8_t* hap, const uint8_t* gap_open,
                      int gap_extend, int nuc_prior, int* n_dp);

/* L2 with the two run-time modes: hap_flank / do_flank = hapFlank / doCalculateFlankScore of
 * calign.pyx:170; hash_read = sequence whose 7-mers vote (NULL = read; HLA mode passes the
 * unclipped read, chaplotype.pyx:637-655). */
int plo_map_and_align_ex(const uint8_t* read, const uint8_t* qual, int read_start, int hap_start,
                         int read_len, int hap_len, const uint8_t* hap, const uint8_t* gap_open,
                         int gap_extend, int nuc_prior, int hap_flank, int do_flank,
                         const uint8_t* hash_read, int hash_read_len, int* n_dp);

/* src/cython/chaplotype.pyx:621-676, default mode (useMapQualCap = 0). */
double plo_score_to_ll(int score, int mapq);
/* same, useMapQualCap = 1 (--HLATyping=1): cap = mLTOT*mapq, smooth flattening above score 100. */
double plo_score_to_ll_hla(int score, int mapq);

/* src/cython/chaplotype.pyx:103-115 */
int plo_overlap(int hap_start, int hap_end, int read_pos, int read_end);

/* S2 for a whole batch: fills out->ll / out->score (layout of PlbLoglikOut).
 * n_threads > 1 shards windows over OpenMP threads the way runner.py:470-474 deals
 * regions to processes.  stats may be NULL. */
int plo_window_loglik(const PlbWindowBatch* b, const PlbOptions* opt, PlbLoglikOut* out,
                      int n_threads, PlbRunStats* stats);

/* S3 for a whole batch: cgenotype.pyx:131-189, cpopulation.pyx:268-309, 384-457,
 * 459-594, 623-720.  ll may be NULL (scratch is allocated). */
int plo_population_run(const PlbWindowBatch* b, const PlbOptions* opt, PlbPopulationOut* out,
                       PlbLoglikOut* ll, int n_threads, PlbRunStats* stats);

/* N4: computeGenotypeCallAndLikelihoods + the per-sample derivations of outputCallToVCF
 * (src/cython/vcfutils.pyx:163-334, 491-548) for every (site, individual). */
int plo_site_genotypes(const PlbWindowBatch* b, const PlbPopulationOut* pop, const PlbSiteBatch* sites,
                       PlbSiteOut* out);

/* Pieces of S3 exposed for unit tests. */
double plo_genotype_loglik(const double* ll1, const double* ll2, int n_total, int n_good,
                           int homozygous, double* gof, double* hap1_like, double* hap2_like);
int plo_em(const double* gl, const int32_t* n_reads, int n_ind, int n_hap, int gl_stride,
           int max_iters, double* freq, double* em_post);
double plo_posterior(const double* gl, const int32_t* n_reads, int n_ind, int n_hap, int gl_stride,
                     const double* freq, uint64_t const* hap_var_mask, int var, double prior);

#ifdef __cplusplus
}
#endif
#endif
