/*
 * platypus_b200.h — C ABI of the B200 read-vs-haplotype likelihood engine.
 *
 * This is the drop-in boundary for ONE path of andyrimmer/Platypus: scoring reads
 * against candidate haplotypes (banded affine-gap min-plus alignment + 7-mer anchor
 * voting), turning scores into per-read log-likelihoods, and reducing those to
 * genotype likelihoods, EM haplotype frequencies and variant posteriors.
 *
 * Every entry point below names the reference interface it replaces (paths relative
 * to the reference checkout).  Plain pointers and sizes only; the caller owns every
 * buffer; every function returns an int status (0 = PLB_OK, negative = error, text
 * via plb_last_error()).  No entry point ever falls back to a CPU implementation:
 * if no CUDA device / kernel image is available the call fails with PLB_ERR_CUDA.
 */
#ifndef PLATYPUS_B200_H
#define PLATYPUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PLB_ABI_VERSION 6   /* 6: plb_best_score_genotypes_host */

/* status codes */
#define PLB_OK              0
#define PLB_ERR_ARG        -1   /* NULL pointer / inconsistent offsets                         */
#define PLB_ERR_SHAPE      -2   /* limit exceeded (haplotype > 16384 bp, H > max_haps, ...)     */
#define PLB_ERR_CUDA       -3   /* CUDA runtime error, no device, or kernel image missing       */
#define PLB_ERR_UNSUPPORTED -4  /* entry point does not take this input format (2-bit packed batches: only the S2 / S3
                                   window calls and the device-resident path read them)              */
#define PLB_ERR_NOMEM      -5

/* limits inherited from the reference */
#define PLB_MAX_HAP_LEN    16384  /* hash_size, src/cython/calign.pyx:25-27, chaplotype.pyx:180-183 */
#define PLB_KMER           7      /* hash_nucs,  src/cython/calign.pyx:25                           */
#define PLB_BAND           16     /* diagonals,  src/c/align.c:83-88                                */
#define PLB_SCORE_NONE     1000000 /* "no alignment attempted" sentinel, calign.pyx:184            */
#define PLB_LL_CAP         (-300.0) /* chaplotype.pyx:634                                           */

/*
 * Scoring options.  Mirrors the constants hard-wired in
 * src/cython/chaplotype.pyx:606-608 (gapExtend=3, nucprior=2) and the run-time
 * switches the hot path reads from `options`: HLATyping,
 * calculateFlankScore, useEMLikelihoods; max_em_iters is the `maxIters` argument of
 * Population.call (src/cython/variantcaller.pyx:141 passes 100).
 */
typedef struct PlbOptions {
    int32_t gap_extend;         /* 3 */
    int32_t nuc_prior;          /* 2 */
    int32_t use_mapq_cap;       /* options.HLATyping = useMapQualCap of alignReads (cpopulation.pyx:296):
                                   reads clipped to the haplotype, cap = mLTOT*mapq, smooth cap above
                                   score 100 (chaplotype.pyx:631-672); 0 or 1                          */
    int32_t calc_flank_score;   /* options.calculateFlankScore: every band alignment with a positive
                                   score loses its in-flank cost (calign.pyx:236-238, 262-264,
                                   align.c:593-644); flank = win_start - hap_start must be > 0; 0 or 1 */
    int32_t max_em_iters;       /* 100 */
    int32_t use_em_likelihoods; /* options.useEMLikelihoods, cpopulation.pyx:654-657            */
} PlbOptions;

/*
 * A batch of windows.  One window = what one call of callVariantsInWindow
 * (src/cython/variantcaller.pyx:74-141) hands to Population.setup: a list of
 * haplotypes sharing one [win_start, win_end) interval and one flank, and for every
 * individual three read lists (good, bad, broken mates) in that order
 * (src/cython/chaplotype.pyx:341-373).
 *
 * Reads live in a pool and windows refer to them by index (slot_read), which is the
 * flat equivalent of the reference's cAlignedRead** window pointers
 * (src/cython/cwindow.pyx:208-236): a read overlapping several windows is stored and
 * copied to the GPU once.  Field meaning follows cAlignedRead
 * (src/cython/htslibWrapper.pxd:187-201): seq = ASCII bases, qual = raw phred (no
 * +33), pos/end = alignment start/end, mapq, QC-fail = bitFlag & 512.
 *
 * All pointers are HOST pointers for the *_host entry points and DEVICE pointers for
 * the *_device entry points.
 */
typedef struct PlbWindowBatch {
    int32_t n_windows;
    int32_t n_individuals;      /* nInd, identical for all windows of the batch            */
    int32_t n_haps;             /* total haplotypes = win_hap_off[n_windows]                */
    int32_t n_reads;            /* reads in the pool                                        */
    int64_t n_slots;            /* window-read slots = wi_slot_off[n_windows*n_individuals] */

    /* per window */
    const int32_t* win_hap_off; /* [n_windows+1] haplotype index range of each window       */
    const int32_t* win_start;   /* [n_windows] Haplotype.startPos                           */
    const int32_t* win_end;     /* [n_windows] Haplotype.endPos                             */
    const int32_t* hap_start;   /* [n_windows] genomic position of base 0 of the window's
                                   haplotype sequences = startPos - endBufferSize
                                   (src/cython/chaplotype.pyx:604)                          */

    /* per haplotype */
    const int64_t* hap_seq_off; /* [n_haps+1] byte offsets into hap_seq                     */
    const uint8_t* hap_seq;     /* ASCII, Haplotype.cHaplotypeSequence (flank+window+flank) */

    /* per (window, individual), index w*n_individuals + i */
    const int64_t* wi_slot_off; /* [n_windows*n_individuals+1] slot range, ordered good|bad|broken */
    const int32_t* wi_n_good;   /* [n_windows*n_individuals] reads.windowEnd-windowStart    */
    const int32_t* wi_n_bad;    /* [n_windows*n_individuals] badReads count                 */

    /* per slot */
    const int32_t* slot_read;   /* [n_slots] index into the read pool                       */

    /* read pool */
    const int64_t* read_seq_off;/* [n_reads+1] byte offsets into read_seq / read_qual       */
    const uint8_t* read_seq;
    const uint8_t* read_qual;
    const int32_t* read_pos;    /* [n_reads] cAlignedRead.pos                               */
    const int32_t* read_end;    /* [n_reads] cAlignedRead.end                               */
    const uint8_t* read_mapq;   /* [n_reads]                                                */
    const uint8_t* read_qcfail; /* [n_reads] nonzero = Read_IsQCFail                        */

    /* variants, only read by plb_population_*; may be NULL / 0 when posteriors are not wanted */
    int32_t max_variants;       /* stride of var_prior / var_phred, <= 64                   */
    const int32_t* win_n_var;   /* [n_windows]                                              */
    const uint64_t* hap_var_mask;/* [n_haps] bit v set iff window variant v is in hap.variants
                                    (the `var not in vsf` test, cpopulation.pyx:509-516)    */
    const double* var_prior;    /* [n_windows*max_variants] Variant.calculatePrior value    */

    /* Packed bases (ABI 4; zero = the byte-per-base layout above).  With seq_format = PLB_SEQ_2BIT hap_seq and read_seq
     * hold 2 bits per base - A 0, C 1, G 2, T 3; base i of the (concatenated) array sits at bits 2*(i & 3) of byte
     * i >> 2 - which is what BAM's 4-bit nibbles (src/cython/htslibWrapper.pyx:414-416) pack into without ever
     * becoming ASCII.  hap_seq_off / read_seq_off are unchanged: they count BASES (read_qual keeps one byte per base at
     * the same offsets).  Every base that is not exactly 'A', 'C', 'G' or 'T' (N, IUPAC codes, lower case) is listed
     * with its base index and its original byte; the packed code at that position is ignored.  Results are
     * bit-identical to the ASCII call: the bytes are restored on the GPU before anything reads them.
     * plb_pack_bases_host / plb_pack_nibbles_host produce this layout. */
    int32_t seq_format;
    int64_t n_read_exc;          /* exceptions of read_seq, ascending positions             */
    const int64_t* read_exc_pos; /* [n_read_exc] base index into the read pool               */
    const uint8_t* read_exc_chr; /* [n_read_exc] the byte the reference would see           */
    int64_t n_hap_exc;
    const int64_t* hap_exc_pos;
    const uint8_t* hap_exc_chr;

    /* Packed qualities (ABI 5; zero = one byte per base).  qual_bits = 4 or 6: read_qual holds qual_bits-bit codes into
     * qual_table, the code of base i of the read pool at bit i * qual_bits of the byte stream (bit b = bit b & 7 of byte
     * b >> 3).  Base qualities take few distinct values - at most 42 with Illumina's scale, a handful with binned
     * qualities - so this is lossless whenever a batch has <= 64 (<= 16) of them; plb_pack_quals_host builds table and
     * codes, or reports that the batch has to stay at 8 bits.  Offsets are unchanged (they count bases); results are
     * bit-identical: the bytes are restored on the GPU before anything reads them. */
    int32_t qual_bits;
    uint8_t qual_table[64];
} PlbWindowBatch;

#define PLB_SEQ_ASCII 0
#define PLB_SEQ_2BIT  1

/*
 * Per-read outputs of the scoring stage (seam S2): what
 * Haplotype.alignReads (src/cython/chaplotype.pyx:306-377) leaves in
 * likelihoodCache, for every haplotype of every window at once.
 * Layout: for (w,i) the block starts at ll_off[w*nInd+i] and is [H_w][T_wi]
 * (haplotype-major, reads in good|bad|broken order); no 999 sentinel.
 */
typedef struct PlbLoglikOut {
    const int64_t* ll_off;      /* [n_windows*n_individuals+1], see plb_ll_offsets          */
    double*  ll;                /* log-likelihoods; may be NULL                             */
    int32_t* score;             /* best integer alignment score of mapAndAlignReadToHaplotype
                                   (src/cython/calign.pyx:170-272); -1 where the read was
                                   short-circuited to LL=0 (QC fail / overlap<7); may be NULL */
} PlbLoglikOut;

/*
 * Per-window outputs of the population model (seam S3): the fields
 * variantcaller.pyx:582 reads back from Population after setup()+call()
 * (src/cython/cpopulation.pxd:15-35).  Fixed strides: Hmax = max_haps,
 * Gmax = Hmax*(Hmax+1)/2, genotype order (i,j), i<=j, row-major
 * (src/cython/cgenotype.pyx:193-218).  Any pointer may be NULL to skip that output.
 */
typedef struct PlbPopulationOut {
    int32_t  max_haps;
    double*  gl;                /* [W][nInd][Gmax] genotypeLikelihoods (rescaled, max = 1)  */
    double*  gl_log_max;        /* [W][nInd]       maxLogLikelihoods                        */
    double*  gof;               /* [W][Gmax][nInd] goodnessOfFitValues                      */
    double*  hap_like;          /* [W][nInd][Hmax] sum_r log10(e)*LL  (DiploidGenotype.hap1Like) */
    double*  freq;              /* [W][Hmax]       frequencies after EM                     */
    double*  em_post;           /* [W][nInd][Gmax] EMLikelihoods                            */
    int32_t* call;              /* [W][nInd]       index of called genotype, -1 = no reads  */
    double*  var_phred;         /* [W][max_variants] calculatePosterior (phred, rounded)    */
    int32_t* em_iters;          /* [W]             EM iterations performed                  */
} PlbPopulationOut;

typedef struct PlbContext PlbContext;

/* -- context -------------------------------------------------------------------- */

/* One context per GPU (and per host thread driving it).  `device` is the CUDA ordinal.
 * `stream` is a cudaStream_t passed as void* (NULL = the context creates its own); with
 * PyTorch pass torch.cuda.current_stream().cuda_stream so events recorded by torch see
 * the kernels. */
int  plb_context_create(int device, void* stream, PlbContext** out);
void plb_context_destroy(PlbContext* ctx);
const char* plb_last_error(void);
int  plb_abi_version(void);
/* number of kernel launches issued through this context since creation */
int64_t plb_launch_count(const PlbContext* ctx);

/* -- helpers (host only, no GPU work) ---------------------------------------------- */

/* Fills ll_off[n_windows*n_individuals+1] for PlbLoglikOut from the HOST batch; returns
 * total element count in *total. */
int plb_ll_offsets(const PlbWindowBatch* host_batch, int64_t* ll_off, int64_t* total);
/* Validates a HOST batch against the reference's limits (haplotype <= 16384 bp,
 * hapLen >= readLen+15 for every scored read, H <= max_haps ...). */
int plb_validate(const PlbWindowBatch* host_batch, const PlbOptions* opt, int32_t max_haps);

/* -- S1: kernel seam ------------------------------------------------------------- */

/*
 * Replaces fastAlignmentRoutine (src/c/align.h:8-9, src/c/align.c:77-586).  Same argument
 * meaning: seq1 = haplotype segment of len1 = len2+15 bases, seq2/qual2 = read,
 * localgapopen >= len1 entries.  As in the reference, traceback is requested by passing
 * aln1/aln2 (both, 2*len2+16 bytes each, align.c:96): they receive the NUL-terminated
 * alignment rows and *firstpos the segment column of the first row (align.c:523-577,
 * same tie-breaking: M before I before D).  Host buffers; launches a 1-element batch.
 * Returns the score (>= 0) or a negative status.
 */
int plb_fast_align(PlbContext* ctx, const char* seq1, const char* seq2, const char* qual2,
                   int len1, int len2, int gapextend, int nucprior,
                   const char* localgapopen, char* aln1, char* aln2, int* firstpos);

/*
 * Batched S1: n explicit (read, haplotype segment) alignments.  Host buffers.
 * hap_seg_off/read_off are [n+1] byte offsets; segment i must hold
 * (read_len_i + 15) bases and gap_open the same number of entries.
 */
int plb_align_batch_host(PlbContext* ctx, int32_t n,
                         const int64_t* hap_seg_off, const uint8_t* hap_seg, const uint8_t* gap_open,
                         const int64_t* read_off, const uint8_t* read_seq, const uint8_t* read_qual,
                         int gapextend, int nucprior, int32_t* scores_out);

/*
 * Batched fastAlignmentRoutine WITH traceback.  Alignment i writes its rows at
 * aln1_out/aln2_out + 2*read_off[i] + 16*i (2*len+16 bytes, NUL-terminated) and
 * firstpos_out[i].
 */
int plb_align_traceback_host(PlbContext* ctx, int32_t n,
                             const int64_t* hap_seg_off, const uint8_t* hap_seg, const uint8_t* gap_open,
                             const int64_t* read_off, const uint8_t* read_seq, const uint8_t* read_qual,
                             int gapextend, int nucprior, int32_t* scores_out,
                             char* aln1_out, char* aln2_out, int32_t* firstpos_out);

/*
 * Batched fastAlignmentRoutine + calculateFlankScore (src/c/align.h:11-12, align.c:593-644)
 * as mapAndAlignReadToHaplotype chains them (calign.pyx:232-238): alignment i runs on the
 * band segment that starts at seg_start[i] of WHOLE haplotype i (hap_off/hap_seq/gap_open
 * index whole haplotypes), flank_out[i] = calculateFlankScore(hapLen_i, hap_flank[i], quals,
 * localgapopen, gapextend, nucprior, firstpos + seg_start[i], aln1, aln2) of the alignment the
 * traceback produces.  Computed in one forward pass (no back-pointer storage).
 */
int plb_align_flank_batch_host(PlbContext* ctx, int32_t n,
                               const int64_t* hap_off, const uint8_t* hap_seq, const uint8_t* gap_open,
                               const int32_t* seg_start, const int32_t* hap_flank,
                               const int64_t* read_off, const uint8_t* read_seq, const uint8_t* read_qual,
                               int gapextend, int nucprior, int32_t* scores_out, int32_t* flank_out);

/*
 * Replaces Haplotype.annotateWithGapOpen (src/cython/chaplotype.pyx:552-590) for a
 * set of haplotypes.  out must hold hap_seq_off[n_haps] + n_haps bytes: haplotype h
 * gets hapLen_h+1 entries starting at hap_seq_off[h] + h (the last one is 0).
 */
int plb_gap_open_host(PlbContext* ctx, int32_t n_haps, const int64_t* hap_seq_off,
                      const uint8_t* hap_seq, uint8_t* out);

/* -- packing helpers (host only) ---------------------------------------------------------- */

/*
 * Packs n ASCII bases (src) as bases dst_base .. dst_base+n-1 of the 2-bit array dst (bits are OR-ed in: dst must be
 * zero-initialised; any dst_base, so reads can be appended one at a time) and appends every base that is not exactly
 * A/C/G/T to exc_pos / exc_chr as (dst_base + i, byte), starting at *n_exc; PLB_ERR_SHAPE when more than exc_cap entries
 * would be needed.  This is the staging step of row N3 (BAM record -> device read pool) for callers that hold ASCII
 * reads, as Platypus does after ReadIterator.get (src/cython/htslibWrapper.pyx:328-406).
 */
int plb_pack_bases_host(const uint8_t* src, int64_t n, uint8_t* dst, int64_t dst_base,
                        int64_t* exc_pos, uint8_t* exc_chr, int64_t exc_cap, int64_t* n_exc);
/*
 * Same from BAM's own 4-bit encoding (two bases per byte, high nibble first, codes "=ACMGRSVTWYHKDBN",
 * htslibWrapper.pyx:414-416): the read never exists as ASCII on the host.  Exceptions carry the letter htslib's
 * lookup table gives the nibble.
 */
int plb_pack_nibbles_host(const uint8_t* bam_seq, int64_t n, uint8_t* dst, int64_t dst_base,
                          int64_t* exc_pos, uint8_t* exc_chr, int64_t exc_cap, int64_t* n_exc);

/*
 * Packs n raw base qualities into 4- or 6-bit codes (see PlbWindowBatch.qual_bits): qual_table receives the distinct
 * values in ascending order, *qual_bits 4 when there are <= 16 of them, 6 when <= 64; dst needs (n * 6 + 7) / 8 + 4 bytes.
 * More than 64 distinct values, or a value above 93: PLB_ERR_SHAPE and nothing is written (send the bytes as they are).
 */
int plb_pack_quals_host(const uint8_t* src, int64_t n, uint8_t* dst, int32_t* qual_bits, uint8_t* qual_table);

/* -- S2: per-read scoring seam ----------------------------------------------------- */

/*
 * Replaces, for every (window, individual, haplotype) of the batch at once,
 * Haplotype.alignReads / alignSingleRead (src/cython/chaplotype.pxd:44-45,
 * chaplotype.pyx:306-384) and below them mapAndAlignReadToHaplotype
 * (src/cython/calign.pxd:11) and alignReadToHaplotype (chaplotype.pyx:594-676).
 */
int plb_window_loglik_host(PlbContext* ctx, const PlbWindowBatch* host_batch,
                           const PlbOptions* opt, PlbLoglikOut* host_out);

/* -- S3: window model seam --------------------------------------------------------- */

/*
 * Replaces Population.setup + Population.call (src/cython/cpopulation.pxd:46-56,
 * cpopulation.pyx:197-309, 678-720) for every window of the batch.  Host buffers in,
 * host buffers out; H2D/D2H copies are part of the call.  `ll` may be NULL or a
 * PlbLoglikOut whose buffers also receive the per-read values.
 */
int plb_population_run_host(PlbContext* ctx, const PlbWindowBatch* host_batch,
                            const PlbOptions* opt, PlbPopulationOut* host_out,
                            PlbLoglikOut* host_ll);

/*
 * The same call split in two so that a region loop (src/cython/variantcaller.pyx:566-615) can keep the GPU and the PCIe
 * link busy at once: plb_population_submit queues the uploads, kernels and downloads of one batch and returns without
 * waiting; plb_population_wait blocks until that batch's outputs are in the caller's buffers (and returns its status).
 * Up to PLB_MAX_JOBS batches may be in flight per context - batch k+1 travels over PCIe while batch k computes; a further
 * submit returns PLB_ERR_ARG.  Input and output buffers must stay valid and untouched until the job has been waited
 * for; they should be pinned (cudaHostAlloc / cudaHostRegister), pageable buffers make the copies synchronous.  Jobs
 * complete in submission order.  plb_population_run_host == submit + wait.  `host_out` may be NULL (per-read outputs
 * only: the S2 call).  Any OUTPUT pointer may also be a device pointer of the context's GPU (unified addressing): that
 * block then stays in HBM - e.g. the genotype likelihoods a multi-GPU caller all-gathers next.
 */
#define PLB_MAX_JOBS 2
typedef struct PlbJob PlbJob;
int plb_population_submit(PlbContext* ctx, const PlbWindowBatch* host_batch, const PlbOptions* opt,
                          PlbPopulationOut* host_out, PlbLoglikOut* host_ll, PlbJob** job);
int plb_population_wait(PlbContext* ctx, PlbJob* job);

/* -- N4: per-site genotype calls (the step after the window model) --------------------- */

/*
 * Sites of a batch of windows, as outputCallToVCF walks them (src/cython/vcfutils.pyx:391-417):
 * one site = one reported position with the variants reported there (variantsThisPos).  Variants are
 * named by their window-local index (the bit of hap_var_mask); hap_is_ref is
 * haplotypeIsRefAtThisPos (vcfutils.pyx:403-417: 0 where a variant of the haplotype spans the
 * position), one byte per haplotype of the site's window.
 */
typedef struct PlbSiteBatch {
    int32_t n_sites;
    const int32_t* site_win;      /* [n_sites] window of the site                                   */
    const int32_t* site_var_off;  /* [n_sites+1] offsets into site_var                               */
    const int32_t* site_var;      /* window-local variant indices, variantsThisPos order             */
    const int64_t* site_hap_off;  /* [n_sites+1] offsets into hap_is_ref (H_w entries per site)       */
    const uint8_t* hap_is_ref;
    int32_t min_posterior;        /* options.minPosterior (runner.py: 5)                              */
} PlbSiteBatch;

/*
 * Per (site, individual): what computeGenotypeCallAndLikelihoods (src/cython/vcfutils.pyx:163-334)
 * returns and what outputCallToVCF derives from it (vcfutils.pyx:504-548).  Allele pairs are ordered
 * as the reference loops them: (0,0),(1,0),(1,1),(2,0),(2,1),(2,2)...; stride max_pairs.
 * Any pointer may be NULL.  Individuals without reads get GT -1/-1 and zeros (vcfutils.pyx:497-499).
 * The final text formatting (round(log10 GL, 2), int(GOF), the minReads rule on per-variant read
 * counts) is the VCF writer's and is not done here.
 */
typedef struct PlbSiteOut {
    int32_t  max_pairs;
    int32_t* phased;      /* [n_sites][nInd][2]  phasedIndex1, phasedIndex2                          */
    double*  lik;         /* [n_sites][nInd][max_pairs] marginal genotype likelihoods                 */
    double*  post;        /* [n_sites][nInd][3]  genotype, non-ref and ref posterior                  */
    int32_t* phred;       /* [n_sites][nInd][3]  GQ = phredPosterior, phredNonRef, phredRef           */
    double*  gof;         /* [n_sites][nInd]     bestGoodnessOfFitValue                               */
    int32_t* gt;          /* [n_sites][nInd][2]  GT after the minPosterior rules, -1 = "."            */
    double*  gl_log10;    /* [n_sites][nInd][3]  log10(max(L/maxL, 1e-300)); -1 at multi-allelic sites */
} PlbSiteOut;

/*
 * Replaces the per-sample loop of outputCallToVCF around computeGenotypeCallAndLikelihoods for every
 * site of a batch.  Host buffers.  `batch` supplies win_hap_off, hap_var_mask (varInHap) and wi_n_good;
 * `pop` the gl / gof / freq arrays a previous plb_population_run_host produced for the same batch.
 */
int plb_site_genotypes_host(PlbContext* ctx, const PlbWindowBatch* batch, const PlbPopulationOut* pop,
                            const PlbSiteBatch* sites, PlbSiteOut* out);

/* -- N1: haplotype construction and the haplotype selection loop (the step before the window model) -- */

/*
 * The candidate variants of a batch of windows: the fields of Variant (src/cython/variant.pyx:109-145) that
 * Haplotype.getMutatedSequence (src/cython/chaplotype.pyx:397-449), isHaplotypeValid
 * (src/cython/platypusutils.pyx:735-802) and Variant.__richcmp__ (variant.pyx:282-353) read.  Per window in the
 * order of the window's `variants` list, which the reference keeps sorted (position, type, nRemoved) and free of
 * duplicates; both are checked (PLB_ERR_ARG).  At most 64 variants per window (bit v of a haplotype mask = variant v).
 */
typedef struct PlbVariantSet {
    const int32_t* win_var_off;    /* [n_windows+1] variant index range of each window                  */
    const int32_t* var_pos;        /* Variant.refPos                                                     */
    const int32_t* var_n_removed;  /* len(Variant.removed)                                               */
    const int32_t* var_n_support;  /* Variant.nSupportingReads; only read by plb_select_haplotypes_host  */
    const int64_t* var_added_off;  /* [n_vars+1] byte offsets into var_added                             */
    const uint8_t* var_added;      /* Variant.added, ASCII                                               */
} PlbVariantSet;

/* The options getFilteredHaplotypes reads (src/python/runner.py:519-597, variantcaller.pyx:920). */
typedef struct PlbSelectOptions {
    int32_t max_haplotypes;           /* options.maxHaplotypes (50)                                      */
    int32_t original_max_haplotypes;  /* options.originalMaxHaplotypes (= maxHaplotypes); the heap holds
                                         original_max_haplotypes - 1 <= 63 entries                        */
    int32_t max_variants;             /* options.maxVariants (8)                                         */
    int32_t filter_vars_by_coverage;  /* options.filterVarsByCoverage (1)                                */
    int32_t coverage_sampling_level;  /* options.coverageSamplingLevel (30), > 0                         */
} PlbSelectOptions;

/* Per window: the variant sets of the haplotypes getFilteredHaplotypes returns, in its order, as bit masks
 * over the window's variants; the reference haplotype is NOT in the list (the caller adds it, as
 * variantcaller.pyx:116-120 does). */
typedef struct PlbSelectOut {
    int32_t   max_sel;    /* stride of sel_mask / sel_score; a window with more haplotypes -> PLB_ERR_SHAPE */
    int32_t*  n_sel;      /* [W]                                                                          */
    uint64_t* sel_mask;   /* [W][max_sel]                                                                 */
    double*   sel_score;  /* [W][max_sel] computeBestScoreForGenotype of the kept haplotype; NaN in the
                             enumerate-all branch (variantFilter.pyx:411-438); may be NULL                 */
    int32_t*  n_scored;   /* [W] trial haplotypes scored (nHapsDone); may be NULL                          */
} PlbSelectOut;

/*
 * Replaces Haplotype.__init__'s sequence construction (src/cython/chaplotype.pyx:127-172, 397-449) for n_haps
 * haplotypes at once: haplotype k = the reference haplotype of window hap_win[k] (`ref_batch` holds exactly ONE
 * haplotype per window: Haplotype.referenceSequence = refFile.getSequence(startPos - endBufferSize, endPos +
 * endBufferSize), with win_start / win_end / hap_start as everywhere) mutated by the variants in hap_mask[k]
 * (mask 0 = the reference haplotype itself).  Variants must lie inside [win_start, win_end].  Only the window
 * fields and the haplotype arrays of ref_batch are read.  hap_seq_off[n_haps+1] is always written (lengths are
 * computed on the host); the sequences are built on the GPU and written to hap_seq unless it is NULL (ctx may then be
 * NULL as well: no GPU work); `capacity` = bytes available in hap_seq.
 */
int plb_build_haplotypes_host(PlbContext* ctx, const PlbWindowBatch* ref_batch, const PlbVariantSet* vars,
                              int32_t n_haps, const int32_t* hap_win, const uint64_t* hap_mask,
                              int64_t* hap_seq_off, uint8_t* hap_seq, int64_t capacity);

/*
 * Replaces getFilteredHaplotypes (src/cython/variantFilter.pyx:377-506) with computeBestScoreForGenotype
 * (:237-283) for every window of a batch.  `ref_batch`: ONE haplotype per window (the refHaplotype argument) and,
 * per (window, individual), the good-read list reads.windowStart..windowEnd (wi_n_good; bad reads and broken
 * mates are ignored, as the reference ignores them here).  windowSize = win_end - win_start.
 * Windows with few variants return every valid combination (no GPU work); the others run the reference's rounds
 * - one per variant in order of decreasing nSupportingReads - batched over all windows: per round the trial
 * haplotypes of every window are built on the GPU, every sampled read is scored against them (the S2 kernels;
 * Haplotype.alignSingleRead), sum_r log(0.5 (e^LL_ref + e^LL_trial)) is reduced per (trial, individual) on the
 * GPU, and the host replays the reference's heap operations (heapq / sorted / tuple comparison of
 * (score, variants), including their behaviour on tied scores) on the returned scores.
 */
int plb_select_haplotypes_host(PlbContext* ctx, const PlbWindowBatch* ref_batch, const PlbVariantSet* vars,
                               const PlbSelectOptions* sel, const PlbOptions* opt, PlbSelectOut* out);

/*
 * Replaces computeBestScoreForHaplotype (src/cython/variantFilter.pyx:212-234; used by getAllHLAHaplotypesInRegion, :698-706)
 * for every haplotype of a batch: per individual the sum of Haplotype.alignSingleRead(read, False) over the good reads
 * reads.windowStart..windowEnd (bad reads and broken mates are ignored), best individual; an individual without reads
 * sums to 0.0, as in the reference.  score_out[n_haps].
 */
int plb_best_score_haplotypes_host(PlbContext* ctx, const PlbWindowBatch* batch, const PlbOptions* opt, double* score_out);

/*
 * Replaces computeBestScoreForGenotype (src/cython/variantFilter.pyx:237-283) for arbitrary pairs of haplotypes of a batch -
 * the second pass of getAllHLAHaplotypesInRegion scores every haplotype as a genotype with the best one (:723-733):
 * per individual the sum of log(0.5 (exp(s1) + exp(s2))) over every sampleRate-th good read, s = alignSingleRead(read, False),
 * sampleRate = max(1, (rlen of the first read * nReads // (win_end - win_start)) // target_coverage); best individual
 * (one without reads is skipped; -1e20 when none has any).  hap1[k], hap2[k]: haplotype indices of the batch, both in
 * one window.  score_out[n_pairs].  The --HLATyping loop itself (heap of (score, haplotype) tuples, the two passes, the
 * output order) is host bookkeeping on top of this call and plb_best_score_haplotypes_host:
 * platypus_b200/compat.py get_all_hla_haplotypes.
 */
int plb_best_score_genotypes_host(PlbContext* ctx, const PlbWindowBatch* batch, const PlbOptions* opt, int32_t target_coverage,
                                  int32_t n_pairs, const int32_t* hap1, const int32_t* hap2, double* score_out);

/*
 * The host-side bookkeeping of plb_select_haplotypes_host alone (trial sets per round, isHaplotypeValid, the heap /
 * sort replay, the final ranking), with the score of every trial haplotype supplied by the caller: per round
 * `score` receives the trial haplotypes (window of the caller's batch + variant mask) and fills score_out (nonzero
 * return = abort).  No GPU work and no scoring arithmetic: a hook for callers that score elsewhere and for the CPU
 * tests of the bookkeeping.  Only the window / haplotype arrays of ref_batch are read.
 */
typedef int (*plb_trial_score_fn)(void* user, int32_t n_haps, const int32_t* hap_win, const uint64_t* hap_mask,
                                  double* score_out);
int plb_select_replay_host(const PlbWindowBatch* ref_batch, const PlbVariantSet* vars, const PlbSelectOptions* sel,
                           plb_trial_score_fn score, void* user, PlbSelectOut* out);

/* Device milliseconds of the last plb_select_haplotypes_host on this context, by stage: [0] reference-haplotype
 * pass, [1] haplotype construction, [2] scoring kernels (k_prep..k_dp), [3] score reduction; [4] = host
 * milliseconds in the heap replay and round planning, [5] = round launches (rounds summed over the window groups), [6] = trial haplotypes scored,
 * [7] = (read, haplotype) pairs scored, [8] = algorithmic cells (16 * readLen per pair), [9] = windows that took
 * the scoring rounds.  Writes min(n, 10) values. */
int plb_select_stats(PlbContext* ctx, double* out, int n);

/* -- N3: read staging (the step before the window model) ------------------------------------------------ */

/*
 * n alignment records of a BAM file in file order, as arrays of the record's own fields (SAM specification 4.2):
 * what a BGZF / BAM reader has in hand before anything is decoded.  Sequences stay in BAM's 4-bit encoding.
 */
typedef struct PlbBamRecords {
    int32_t n;
    const int32_t*  ref_id;      /* [n] refID                                                         */
    const int32_t*  pos;         /* [n] 0-based leftmost mapped coordinate                            */
    const uint8_t*  mapq;        /* [n]                                                               */
    const uint16_t* flag;        /* [n] bitwise FLAG                                                  */
    const int32_t*  mate_ref_id; /* [n] next refID                                                    */
    const int32_t*  mate_pos;    /* [n] next pos                                                      */
    const int32_t*  tlen;        /* [n] template length (cAlignedRead.insertSize)                     */
    const int64_t*  cigar_off;   /* [n+1] offsets into cigar                                          */
    const uint32_t* cigar;       /* BAM encoding: op_len << 4 | op                                    */
    const int64_t*  seq_off;     /* [n+1] BASE offsets: l_seq = seq_off[i+1] - seq_off[i]; qual is indexed by them */
    const int64_t*  nib_off;     /* [n+1] BYTE offsets of each record's 4-bit packed sequence in nib   */
    const uint8_t*  nib;         /* two bases per byte, high nibble first, codes "=ACMGRSVTWYHKDBN"    */
    const uint8_t*  qual;        /* raw phred; a record whose first quality byte is 0xFF has none       */
} PlbBamRecords;

/* The read-filter options of bamReadBuffer (src/cython/cwindow.pyx:490-528; defaults of src/python/runner.py:551-580). */
typedef struct PlbReadFilterOptions {
    int32_t min_good_qual_bases;   /* minGoodQualBases 20                 */
    int32_t min_map_qual;          /* minMapQual 20                       */
    int32_t min_base_qual;         /* minBaseQual 20                      */
    int32_t trim_read_flank;       /* trimReadFlank 0                     */
    int32_t trim_overlapping;      /* trimOverlapping 1                   */
    int32_t trim_adapter;          /* trimAdapter 1                       */
    int32_t trim_soft_clipped;     /* trimSoftClipped 1                   */
    int32_t filter_duplicates;     /* filterDuplicates 1                  */
    int32_t filter_mate_unmapped;  /* filterReadsWithUnmappedMates 1      */
    int32_t filter_mate_distant;   /* filterReadsWithDistantMates 1       */
    int32_t filter_small_insert;   /* filterReadPairsWithSmallInserts 1   */
} PlbReadFilterOptions;

/* Outputs of plb_stage_reads_host, caller-allocated, one entry per input record unless noted. */
typedef struct PlbStagedReads {
    uint8_t*  kept;       /* 1 = the record is a read (ReadIterator.get returned one), 0 = no sequence / no qualities */
    uint8_t*  good;       /* 1 = bamReadBuffer.reads, 0 = bamReadBuffer.badReads                       */
    int32_t*  read_pos;   /* cAlignedRead.pos: first base of the read (leading soft clip subtracted)    */
    int32_t*  read_end;   /* cAlignedRead.end: bam_endpos                                              */
    uint16_t* flag_out;   /* bitFlag after Read_SetQCFail                                              */
    uint8_t*  qual_out;   /* [seq_off[n]] qualities after trimming                                     */
    uint8_t*  seq2;       /* [(seq_off[n] + 3) / 4 + 1] bases as 2-bit codes at the records' base offsets: with
                             seq_off as read_seq_off this IS the read_seq of a PLB_SEQ_2BIT batch       */
    int64_t   exc_cap;    /* capacity of exc_pos / exc_chr                                             */
    int64_t   n_exc;      /* out: bases that are not A/C/G/T (ascending)                               */
    int64_t*  exc_pos;
    uint8_t*  exc_chr;
    int32_t   counts[7];  /* filteredReadCountsByType: low-quality bases, unmapped, mate unmapped, mate distant,
                             small insert, duplicate, low mapping quality (cwindow.pyx:40-46); -1 for a
                             filter that is switched off, as in the reference (cwindow.pyx:516-526)     */
} PlbStagedReads;

/*
 * Replaces, for every record of a buffer at once, ReadIterator.get (src/cython/htslibWrapper.pyx:328-406),
 * bamReadBuffer.addReadToBuffer and checkAndTrimRead (src/cython/cwindow.pyx:560-595, 332-481): which list the read
 * joins, its QC-fail flag, its trimmed qualities, its pos / end - and packs the bases from BAM nibbles into the 2-bit
 * read pool of a PLB_SEQ_2BIT batch.  Host only (no GPU work; ctx-free).
 */
int plb_stage_reads_host(const PlbBamRecords* records, const PlbReadFilterOptions* opt, PlbStagedReads* out);

/*
 * Replaces ReadArray.setWindowPointers (src/cython/cwindow.pyx:208-236) for many windows: over a position-sorted read
 * list, window w gets the contiguous slice [lo_out[w], hi_out[w]) from the first read with pos >= max(1, start -
 * longestRead) (skipping leading reads that end at or before start) up to the first read with pos >= end.  The
 * bisection is the reference's own (bisectReadsLeft, cwindow.pyx:276-300), so a list that is slightly out of order - pos
 * has a leading soft clip subtracted, file order does not - yields the reference's slices too.
 */
int plb_window_slices_host(int32_t n_reads, const int32_t* read_pos, const int32_t* read_end, int32_t n_windows,
                           const int32_t* win_start, const int32_t* win_end, int32_t* lo_out, int32_t* hi_out);

/* -- device-resident variants (inputs already in HBM; used by the multi-GPU driver) -- */

/* Copies a HOST batch into device memory owned by the context and returns an opaque
 * handle; the handle stays valid until plb_batch_free. */
typedef struct PlbDeviceBatch PlbDeviceBatch;
int  plb_batch_upload(PlbContext* ctx, const PlbWindowBatch* host_batch, PlbDeviceBatch** out);
void plb_batch_free(PlbContext* ctx, PlbDeviceBatch* b);

/* Runs scoring (+ population model when pop_out != NULL) on a device-resident batch.
 * All pointers inside dev_pop / dev_ll are DEVICE pointers (e.g. torch tensors'
 * data_ptr()).  Asynchronous on the context's stream; no host synchronisation. */
int plb_run_device(PlbContext* ctx, PlbDeviceBatch* batch, const PlbOptions* opt,
                   PlbPopulationOut* dev_pop, PlbLoglikOut* dev_ll);

/* Copies the INPUT arrays of a resident batch back into `host_batch`, which must have the shape (counts and offsets) of
 * the batch that was uploaded: sequences, qualities, read fields, window coordinates, variant masks and priors. */
int plb_batch_download(PlbContext* ctx, PlbDeviceBatch* batch, PlbWindowBatch* host_batch);

/*
 * Measurement support (BASELINE config 5, SURVEY 8d: "generated on device"): overwrites the inputs of a resident ASCII
 * batch in place with the synthetic windows first_window, first_window + 1, ... ("synth-v1d": the recipe of the host
 * generator with a counter-based hash as random source, platypus_b200/csrc/plb_synth.cuh).  Shapes stay those of the
 * uploaded batch - every window's haplotypes of one length, slot s = read s, all reads good - so the launch plan made at
 * upload time serves every refill.  Asynchronous on the context's stream.
 */
int plb_synth_fill_device(PlbContext* ctx, PlbDeviceBatch* batch, uint64_t seed, int64_t first_window);

/* Statistics of the last plb_run_device / *_host call on this context (after a stream
 * synchronise): number of scored (read,haplotype) pairs, band-DP executions, and
 * algorithmic cells = 16 * readLen summed over scored pairs. */
typedef struct PlbRunStats {
    int64_t n_pairs;
    int64_t n_pairs_scored;
    int64_t n_dp;
    int64_t cells;
    int64_t n_anchor_heavy;   /* diagnostic: pairs the first anchor guess left open (second step)  */
    int64_t n_anchor_verify;  /* diagnostic: longest anchor tile: microseconds << 32 | window    */
    int64_t n_anchor_exact;   /* pairs without a strict majority (exact tied-maximum scan)     */
} PlbRunStats;
int plb_last_stats(PlbContext* ctx, PlbRunStats* out);

/* Per-kernel device times, measured with CUDA events recorded on the context's stream around every
 * kernel of plb_run_device / *_host.  plb_set_timing(ctx, 1) starts (and resets) recording;
 * plb_kernel_times synchronises and returns the number of runs averaged (up to the last 64) with
 * the mean milliseconds per run in ms[0..5], order: k_prep, k_anchor, k_general, k_dp, k_genotype,
 * k_population. */
#define PLB_N_KERNELS 6
int plb_set_timing(PlbContext* ctx, int on);
int plb_kernel_times(PlbContext* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* PLATYPUS_B200_H */
