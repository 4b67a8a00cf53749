// plb_kernels.cuh — sm_100a kernels of the read-vs-haplotype likelihood path.
//
//   k_anchor   : per tile (window x slot range x haplotype group): gap-open table, 7-mer index of
//                each haplotype in shared memory, per (read, haplotype) anchor voting, candidate
//                start offsets                     (reference: src/cython/calign.pyx:94-124, 155-165,
//                                                   170-272; src/cython/chaplotype.pyx:552-590)
//   k_general  : band alignments that cannot use the packed path (non-ACGTN haplotype bytes, reads
//                shorter than 9, third and later tied candidates)      (src/c/align.c:77-521)
//   k_dp       : per tile: stage read profiles + haplotype records in shared memory, run one packed
//                band alignment per thread, min per pair, score -> log-likelihood
//                                                  (src/c/align.c:77-521, chaplotype.pyx:306-377, 594-676)
//   k_genotype : per (window, individual): genotype log-likelihoods, goodness of fit, rescale
//                                                  (src/cython/cgenotype.pyx:131-189, cpopulation.pyx:283-309)
//   k_population: per window: EM over haplotype frequencies, genotype calls, variant posteriors
//                                                  (src/cython/cpopulation.pyx:384-457, 459-594, 623-720)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plb_dp.cuh"
#include "plb_kmer.cuh"

namespace plb {

constexpr int kKmer = 7;
constexpr int kHashSize = 16384;
constexpr int kScoreNone = 1000000;
constexpr double kMLTOT = -0.23025850929940459;
// log(1 - exp(mLTOT * mapq)) for every mapping quality (chaplotype.pyx:622), filled by plb_context_create with the HOST's
// libm - the reference's own values on that machine - so that a per-read log-likelihood mLTOT * score + this (two
// separately rounded operations, as the reference's C code performs them: no FMA) is the reference's double bit for bit.
// Sums of them taken in read order are then exact too, which is what keeps exactly tied haplotype scores tied.
__constant__ double c_map_right[256];
constexpr double kLog10E = 0.43429448190325182;
constexpr double kLogHalf = -0.69314718055994529;
constexpr int kRankWords = kHashSize / 32;            // one presence bit per possible 7-mer key
constexpr size_t kRankTabBytes = kRankWords * 8;      // uint2 {bits, number of set bits in earlier words}

// ---- device view of a batch (all pointers are device pointers) --------------------------------
struct DevBatch {
    int32_t n_windows, n_individuals, n_haps, n_reads;
    int64_t n_slots, n_pairs;
    const int32_t* win_hap_off;
    const int32_t* win_start;
    const int32_t* win_end;
    const int32_t* hap_start;
    const int64_t* hap_seq_off;
    const uint8_t* hap_seq;
    const int64_t* wi_slot_off;
    const int32_t* wi_n_good;
    const int32_t* wi_n_bad;
    const int32_t* slot_read;
    const int64_t* read_seq_off;
    const uint8_t* read_seq;
    const uint8_t* read_qual;
    const int32_t* read_pos;
    const int32_t* read_end;
    const uint8_t* read_mapq;
    const uint8_t* read_qcfail;
    int32_t max_variants;
    const int32_t* win_n_var;
    const uint64_t* hap_var_mask;
    const double* var_prior;
    // derived at upload
    const int32_t* slot_wi;    // [n_slots] index w*nInd+i of each slot
    const int32_t* hap_win;    // [n_haps] window of each haplotype
    const int64_t* ll_off;     // [n_windows*nInd+1]
    // scratch
    uint8_t* gap_open;         // haplotype h: hapLen+1 entries at hap_seq_off[h] + h
    uint32_t* win_flags;       // [n_windows] bit0 = haplotype bytes outside ACGTN, bit1 = contains 'N'
    int32_t* cand0;            // [n_pairs] first / second packed-path start offset, -1 = none
    int32_t* cand1;
    int32_t* score;            // [n_pairs] running min over general-path alignments
    uint8_t* read_flags;       // [n_reads] k_read_check: bit 1 = qualities add up beyond the packed recurrence's range
};

struct QualTable {   // value of every packed quality code (PlbWindowBatch.qual_table), passed to kernels by value
    uint32_t w[16];  // as words: a kernel indexes the parameter directly (a byte array cast to words is first copied to
};                   // local memory byte by byte: 64 LDC.U8 + 64 STL.U8 per thread)

struct Tile {
    int32_t w;        // window
    int32_t h0, h1;   // haplotype range (global indices)
    int64_t s0, s1;   // slot range (global indices)
};

struct QueueEntry {
    int64_t pair;
    int32_t hap;
    int32_t slot;
    int32_t start;
    uint32_t clip;    // HLA mode: bases clipped off the read front | clipped read length << 16 (0 = unclipped)
};

struct Queue {
    QueueEntry* e;
    int32_t* count;
    int32_t cap;
};

struct Counters {  // device-side statistics
    unsigned long long n_pairs, n_scored, n_dp, cells;
    unsigned long long n_heavy, n_verify, n_exact;  // n_exact: anchor pairs that needed the exact vote array
    unsigned long long err;   // bit 0: a scored read with readLen + 15 > hapLen; bit 1: a base quality above 93
};
constexpr unsigned long long kErrReadTooLong = 1ull, kErrQuality = 2ull;

struct ScoreParams {
    int32_t ext, nuc;
    int32_t flank;    // options.calculateFlankScore: every band alignment loses its in-flank cost (calign.pyx:236-238)
    int32_t hla;      // options.HLATyping (useMapQualCap): reads clipped to the haplotype, map-qual cap (chaplotype.pyx:631-672)
};

// HLA mode clips a read to the haplotype before scoring (chaplotype.pyx:647-655).  The right-hand
// clip is measured from startPos + hapLen (window start, not haplotype start) exactly as the
// reference does; the 7-mer hashes stay those of the unclipped read (read.hash, :637-638).
struct PairClip {
    int off1;   // bases dropped at the front (readSeq += offset1, readStart += offset1)
    int L;      // clipped read length (may be <= 0)
};
__device__ __forceinline__ PairClip pair_clip(const ScoreParams& sp, int read_pos, int read_len, int hap_start,
                                              int win_start, int hap_len) {
    PairClip c;
    c.off1 = 0;
    c.L = read_len;
    if (sp.hla) {
        int o1 = hap_start - read_pos, o2 = read_pos + read_len - win_start - hap_len;
        o1 = o1 < 0 ? 0 : o1;
        o2 = o2 < 0 ? 0 : o2;
        c.off1 = o1;
        c.L = read_len - o1 - o2;
    }
    return c;
}

// homopolymer gap-open table evaluated from chaplotype.pyx:64-67 (see oracle for the formula)
__constant__ uint8_t c_homopol_q[49] = {45, 42, 41, 39, 37, 32, 28, 23, 20, 19, 17, 16, 15, 14, 13, 12, 11,
                                        11, 10, 9,  9,  8,  8,  7,  7,  7,  6,  6,  6,  5,  5,  5,  4,  4,
                                        4,  3,  3,  3,  3,  2,  2,  2,  2,  2,  1,  1,  1,  1,  1};

// Gap-open penalty of position i (chaplotype.pyx:552-590).  The reference scans right to left
// carrying a run length; equivalently run(i) = number of consecutive positions j > i with
// hap[j] == hap[i], stopping at 'N' (an 'N' never continues a run) and saturating at 48.
__device__ __forceinline__ uint8_t gap_open_at(const uint8_t* hap, int hap_len, int i) {
    const uint8_t c = hap[i];
    int run = 0;
    if (c != 'N') {
        // position i extends the run of i+1 iff hap[i] == hap[i+1] (and that base is not 'N')
        while (run < 48 && i + run + 1 < hap_len && hap[i + run + 1] == c) ++run;
    }
    return c_homopol_q[run];
}

// compare-and-swap of one u16 element of a shared-memory array, through its containing 32-bit word
// (native ATOMS.CAS; the library's 16-bit atomicCAS goes through a generic-address call)
__device__ __forceinline__ unsigned short cas_u16(uint16_t* arr, int idx, unsigned short expect, unsigned short val) {
    u32* word = (u32*)(arr + (idx & ~1));
    const int sh = 16 * (idx & 1);
    u32 cur = *(volatile u32*)word;
    while (true) {
        const unsigned short have = (unsigned short)(cur >> sh);
        if (have != expect) return have;
        const u32 want = (cur & ~(0xFFFFu << sh)) | ((u32)val << sh);
        const u32 old = atomicCAS(word, cur, want);
        if (old == cur) return expect;
        cur = old;
    }
}

// ---------------------------------------------------------------------------------------------
// k_prep: one block per haplotype.  Gap-open table (chaplotype.pyx:552-590) into global scratch and
// the per-window "general path" flag (haplotype bytes outside ACGTN need exact byte compares).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_prep(DevBatch b, int h_base, int ext) {
    const int h = h_base + blockIdx.x;
    const int len = (int)(b.hap_seq_off[h + 1] - b.hap_seq_off[h]);
    const uint8_t* hap = b.hap_seq + b.hap_seq_off[h];
    uint8_t* go = b.gap_open + b.hap_seq_off[h] + h;
    int bad = 0, has_n = 0, small_open = 0;
    for (int i = threadIdx.x; i <= len; i += blockDim.x) {
        if (i < len) {
            const uint8_t g = gap_open_at(hap, len, i);
            go[i] = g;
            small_open |= ((int)g < ext);
            const int c = fast_code(hap[i]);
            bad |= (c == 5);
            has_n |= (c == 4);
        } else {
            go[i] = 0;
        }
    }
    bad = __syncthreads_or(bad);
    has_n = __syncthreads_or(has_n);
    small_open = __syncthreads_or(small_open);
    // bit 0: bytes outside ACGTN (general path); bit 1: contains 'N' (8-op packed variant);
    // bit 2: some gap-open penalty below the gap-extension penalty (6-op instead of 5-op variant)
    if ((bad || has_n || small_open) && threadIdx.x == 0)
        atomicOr((unsigned int*)(b.win_flags + b.hap_win[h]),
                 (bad ? 1u : 0u) | (has_n ? 2u : 0u) | (small_open ? 4u : 0u));
}

// ---------------------------------------------------------------------------------------------
// k_anchor — anchor voting (calign.pyx:206-267) for every (read, haplotype) pair of a tile.
//
// The reference keeps, per haplotype, a 16384-entry hash table plus chains, and per pair a vote
// array of hapLen+readLen counters that it clears, fills and scans.  Here:
//   * one open-addressed table per TILE maps the 14-bit 7-mer hash to a dense id over the union of
//     the group's haplotype 7-mers;
//   * per haplotype, head[id] / next[pos] chains give the positions carrying that 7-mer (the
//     reference's hash_table / next_array, calign.pyx:94-124);
//   * votes are not stored: a Boyer-Moore pass finds the only offset that can hold a strict
//     majority of the votes and a second pass counts it exactly.  A strict majority is the unique
//     maximum, i.e. exactly the single candidate the reference would align (ties impossible).
//   * pairs without a strict majority (split votes, repeats, unrelated reads) go to a per-tile
//     list and are re-voted by a whole warp with a real counter array, reproducing the tied-maximum
//     scan of calign.pyx:222-247.
// ---------------------------------------------------------------------------------------------
struct AnchorPlan {
    const Tile* tiles;
    int32_t n_tiles;
    int32_t max_slots;      // slots per tile upper bound
    int32_t max_group;      // haplotypes per tile upper bound
    int32_t max_pairs;      // slots*haplotypes per tile upper bound
    int32_t next_halfs;     // u16 entries for all next arrays of a group
    int32_t heads_halfs;    // u16 entries of the head area (>= largest union size + 1)
    int32_t mult_halfs;     // u16 entries of the heavy-key bitmap area (kRankWords * 2)
    int32_t rpk_words;      // u32 words for the 2-bit packed reads of a tile
    int32_t hpk_words;      // u32 words for the 2-bit packed haplotypes of a group
    int32_t cnt_words;      // u32 words of one warp's counter array (2 counters per word)
    int32_t n_cnt;          // counter arrays available in shared memory (>= 1)
    // byte offsets of the shared-memory areas (laid out by the host planner, 16-byte aligned)
    uint32_t o_cnt, o_fb, o_ul, o_vl, o_rpk, o_hpk, o_next, o_mult, o_heads, o_slot, o_hmeta, smem_bytes;
};

// host + device: lays out the shared memory of k_anchor from the plan's element counts
inline void anchor_layout(AnchorPlan& ap, size_t slot_bytes) {
    auto al = [](size_t x) { return (x + 15) & ~(size_t)15; };
    size_t o = kRankTabBytes;   // 7-mer key -> dense id: 512 x {presence bits, rank prefix}
    ap.o_fb = (uint32_t)o;     o = al(o + (size_t)ap.max_pairs * 4);
    // The per-warp vote arrays of the exact path share their bytes with the two lists of the decision passes: the exact
    // path runs after the decided pairs have been emitted, when neither list is read any more (a block barrier lies
    // between), and the next haplotype sub-group rewrites both lists from scratch.
    const size_t lists = al((size_t)ap.max_pairs * 4) + al((size_t)ap.max_pairs * 12);
    const size_t votes = al((size_t)ap.n_cnt * ap.cnt_words * 4);
    ap.o_cnt = (uint32_t)o;
    ap.o_ul = (uint32_t)o;
    ap.o_vl = (uint32_t)(o + al((size_t)ap.max_pairs * 4));
    o += lists > votes ? lists : votes;
    ap.o_rpk = (uint32_t)o;    o = al(o + (size_t)ap.rpk_words * 4);
    ap.o_hpk = (uint32_t)o;    o = al(o + (size_t)ap.hpk_words * 4);
    ap.o_next = (uint32_t)o;   o = al(o + (size_t)ap.next_halfs * 2);
    ap.o_mult = (uint32_t)o;   o = al(o + (size_t)ap.mult_halfs * 2);
    ap.o_heads = (uint32_t)o;  o = al(o + (size_t)ap.heads_halfs * 2);
    ap.o_slot = (uint32_t)o;   o = al(o + (size_t)ap.max_slots * slot_bytes);
    ap.o_hmeta = (uint32_t)o;  o = al(o + (size_t)ap.max_group * 12);
    ap.smem_bytes = (uint32_t)o;
}

struct SlotInfo {
    int32_t read;     // read pool index
    int32_t len;      // read length
    int32_t pos;      // read.pos
    int32_t flags;    // bit0 = LL forced to 0 (QC fail / overlap < 7)
    int32_t vub;      // upper bound of the votes this read can cast on any haplotype of the sub-group
    int32_t lph;      // light | heavy << 16: read 7-mers whose id occurs once / more than once in a haplotype
    int32_t poff;     // offset (u32 words) of the 2-bit packed read
    int32_t T;        // reads of this (window, individual): stride between haplotype rows of the LL block
    int64_t pair0;    // LL index of (first haplotype of the window, this slot)
};

__device__ __forceinline__ int read_overlap(int ws, int we, int rp, int re) {  // chaplotype.pyx:103-115
    int s = ws > rp ? ws : rp;
    int e = we < re ? we : re;
    return e > s ? e - s : -1;
}

constexpr int kAnchorThreads = 256;
// k_anchor is latency- and barrier-bound, so resident warps count more than registers: it is compiled for 5 and for 4 CTAs
// per SM (48 / 64 registers per thread) and launch_windows takes the five-CTA build where the tile's shared memory lets
// five fit (BASELINE config 2, measured: 3 CTAs with 80 registers 1.16 ms, 4 x 64 1.03 ms, 5 x 48 0.98 ms, and 0.96 ms
// once the vote arrays shared their bytes with the decision lists; 6 x 40 the same as five).

// One band alignment on the scalar path: any bytes, any length, both run-time modes.  `roff` = bases
// clipped off the read front (HLA mode), L = (clipped) read length.
template <bool kModes>
__device__ __noinline__ int general_dp_now(const DevBatch& b, int h, int read, int start, int roff, int L,
                                           ScoreParams sp) {
    const uint8_t* hapg = b.hap_seq + b.hap_seq_off[h];
    const uint8_t* go = b.gap_open + b.hap_seq_off[h] + h;
    const uint8_t* rs = b.read_seq + b.read_seq_off[read] + roff;
    const uint8_t* rq = b.read_qual + b.read_seq_off[read] + roff;
    if (kModes && sp.flank) {
        const int w = b.hap_win[h];
        const int hap_len = (int)(b.hap_seq_off[h + 1] - b.hap_seq_off[h]);
        // hapFlank = hap.endBufferSize = startPos - hapStart (chaplotype.pyx:604, 609)
        return band_dp_flank_adjusted(hapg, go, rs, rq, L, sp.ext, sp.nuc, start, hap_len,
                                      b.win_start[w] - b.hap_start[w]);
    }
    return band_dp_general(hapg + start, go + start, rs, rq, L, sp.ext, sp.nuc);
}

// Collects the distinct band start offsets of one pair: the first two go to the packed path
// (cand0/cand1), the rest - and everything on the general path - to the queue.
template <bool kModes>
struct Emitter {
    int c0, c1, sc;
    int q0 = -1, q1 = -1, q2 = -1;   // the last three starts sent to the queue (duplicates are common: the
                                     // fallback position usually equals a voted candidate)
    unsigned n_dp;
    __device__ __forceinline__ void emit(const DevBatch& b, const Queue& q, ScoreParams sp, bool slow, int64_t pair,
                                         int h, int64_t gs, int read, int roff, int L, int start) {
        if (start == c0 || start == c1 || start == q0 || start == q1 || start == q2) return;
        ++n_dp;
        if (!slow && c0 < 0) {
            c0 = start;
        } else if (!slow && c1 < 0) {
            c1 = start;
        } else {
            q2 = q1;
            q1 = q0;
            q0 = start;
            const int qi = atomicAdd(q.count, 1);
            if (qi < q.cap) {
                QueueEntry qe;
                qe.pair = pair;
                qe.hap = h;
                qe.slot = (int32_t)gs;
                qe.start = start;
                qe.clip = (kModes && sp.hla) ? ((u32)roff | ((u32)L << 16)) : 0u;
                q.e[qi] = qe;
            } else {  // queue full: run it right here
                const int v = general_dp_now<kModes>(b, h, read, start, roff, L, sp);
                sc = v < sc ? v : sc;
            }
        }
    }
};

// 2-bit codes (calign.pyx:69-74) of the 16 bases seq[i0 .. i0+15], base i0+k at bits 2k; positions
// outside [0, len) give 0.  i0 is a multiple of 16.  Reads five aligned 32-bit words (the arrays are
// padded by 64 bytes) and converts four bases at a time:
//   c = byte & 7;  code = (c & 3) ^ (c == 7)      (A->1, C->3, G->2, T->0, N->2: 7 -> 2)
//   four 2-bit codes of a word gathered into one byte by a multiply (no two partial products overlap)
__device__ __forceinline__ u32 pack16_codes(const uint8_t* __restrict__ seq, int i0, int len) {
    if (i0 < 0 || i0 >= len) return 0u;
    const uintptr_t a = (uintptr_t)(seq + i0);
    const u32* wp = (const u32*)(a & ~(uintptr_t)3);
    const int sh = 8 * (int)(a & 3);
    u32 w[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) w[k] = __ldg(wp + k);
    u32 v = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const u32 x = __funnelshift_r(w[k], w[k + 1], sh) & 0x07070707u;
        const u32 is7 = ((x + 0x01010101u) >> 3) & 0x01010101u;
        const u32 code = (x & 0x03030303u) ^ is7;
        v |= ((code * 0x01041040u) >> 24) << (8 * k);
    }
    const int n = len - i0;
    if (n < 16) v &= (1u << (2 * n)) - 1u;
    return v;
}

// 7-mer key of position p of a packed sequence: its 14 bits, first base in the low bits (a
// digit-reversed copy of the reference's hash - any bijection of the hash gives the same votes)
__device__ __forceinline__ u32 key_at(const u32* pk, int p) {
    return fsr(pk[p >> 4], pk[(p >> 4) + 1], 2 * (p & 15)) & 0x3FFFu;
}

// key -> dense id (1..U) over the union of the group's haplotype 7-mers, 0 when absent.  The table is a
// 16384-bit presence map with a rank directory: id = (set bits before the key) + 1.  One LDS.64, a
// popcount and no probing; built with atomicOr + one 512-entry scan (no CAS loops).
__device__ __forceinline__ u32 tab_lookup(const u32* tab, u32 key) {
    const uint2 e = ((const uint2*)tab)[key >> 5];
    const u32 bit = key & 31u;
    const u32 below = e.x & ((1u << bit) - 1u);
    return ((e.x >> bit) & 1u) ? e.y + (u32)__popc(below) + 1u : 0u;
}

// Decides a pair without a vote array when it can.  Guesses: the offsets implied by the first, the
// last and the middle read 7-mer that occur exactly once in the haplotype.  Each guess is counted
// exactly with count_offset; everything not yet counted is bounded by R = V_ub - (counted votes),
// where V_ub bounds ALL votes the read can cast.  As soon as the best counted offset beats R, no
// uncounted offset can reach it, so the maximum of the reference's vote array (calign.pyx:206-220)
// is attained exactly at the counted offsets with that count - unique or tied.
// res[0..2] = tied-maximum offsets + 1 (kNoCand where unused); returns 1 when decided, else 0.
constexpr int kNoCand = 0x40000000;
constexpr int kPairSkip = 0x40000001;       // nothing to decide (LL forced to 0 / read shorter than 7)
constexpr int kPairUndecided = 0x40000002;  // goes to the exact vote array

struct LightArgs {
    u32 head_off, rpk_off, hpk_off, res_off;
    int nk_read, nk_hap, vub;
    int lp, hh;   // read 7-mers that vote exactly once (light) / possibly several times (heavy)
};

// offset implied by the first read 7-mer in [i0, i1) (walking by step) that occurs exactly once in
// the haplotype; read 7-mer -> id through the tile's table (smem offset 0), id -> position through head
__device__ __forceinline__ int unique_hit_offset(const u32* tab, const uint16_t* head, const u32* rpk, int i0,
                                                 int i1, int step) {
    for (int i = i0; i != i1; i += step) {
        const u32 hd = head[tab_lookup(tab, key_at(rpk, i))];
        if (hd && !(hd & 0x8000u)) return (int)hd - 1 - i;
    }
    return kNoCand;
}

// The decision runs in two steps so that warps stay full.  Step one - every pair, one thread each - tries the single
// guess that settles most pairs (the offset implied by the first unique 7-mer of the read).  The pairs it leaves open
// (votes split by an indel, a first guess on the minority side: about a third of them) are compacted into a list, and
// step two works through that list with the remaining guesses.  Run as one loop per pair, nearly every warp held a lane
// that needed three or four guesses and the other lanes idled through them (21 of 32 threads active per instruction).
//
// light_first: returns 1 when decided (res[0..2] written), else 0 with res[0] = the guess (or kNoCand) and res[1] = its
// count, from which light_rest resumes.
__device__ __noinline__ int light_first(LightArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint16_t* head = (const uint16_t*)(smem + a.head_off);
    const u32* tab = (const u32*)smem;
    const u32* rpk = (const u32*)(smem + a.rpk_off);
    const u32* hpk = (const u32*)(smem + a.hpk_off);
    u32* res = (u32*)(smem + a.res_off);
    const int nk = a.nk_read;
    if (a.vub == 0) {  // no read 7-mer occurs in this haplotype group: maxcount == 0, no candidates
        res[0] = res[1] = res[2] = (u32)kNoCand;
        return 1;
    }
    const int g0 = unique_hit_offset(tab, head, rpk, 0, min(nk, 24), 1);
    int c0 = 0;
    if (g0 != kNoCand) {
        c0 = count_offset_bits(rpk, hpk, nk, a.nk_hap, g0);
        const int R = a.vub - c0, R2 = a.hh + a.lp - max(0, c0 - a.hh);
        if (c0 > min(R, R2) && c0 > 0) {
            res[0] = (u32)(g0 + 1);
            res[1] = res[2] = (u32)kNoCand;
            return 1;
        }
    }
    res[0] = (u32)g0;
    res[1] = (u32)c0;
    return 0;
}

__device__ __noinline__ int count_offset_call(u32 rpk_off, u32 hpk_off, int nk_read, int nk_hap, int idx) {
    extern __shared__ __align__(16) uint8_t smem[];
    return count_offset_bits((const u32*)(smem + rpk_off), (const u32*)(smem + hpk_off), nk_read, nk_hap, idx);
}

// The remaining guesses of a pair light_first left open: read 7-mers near the end, the middle, then the quarters and
// eighths (reads that differ from the haplotype by several indels split their votes over several offsets; each extra
// exact count is ~100x cheaper than the warp-wide vote array).
__device__ __noinline__ int light_rest(LightArgs a) {
    extern __shared__ __align__(16) uint8_t smem[];
    const uint16_t* head = (const uint16_t*)(smem + a.head_off);
    const u32* tab = (const u32*)smem;
    const u32* rpk = (const u32*)(smem + a.rpk_off);
    u32* res = (u32*)(smem + a.res_off);
    const int nk = a.nk_read;
    const int lim = min(nk, 24);
    constexpr int NG = 7;
    int g[NG], c[NG];
    g[0] = (int)res[0];
    c[0] = (int)res[1];
    // Two bounds on the votes of any offset not counted yet:
    //   R  = (all votes the read can cast) - (votes counted so far)
    //   R2 = heavy + (light - light votes counted so far): every read 7-mer gives an offset at most ONE vote,
    //        light 7-mers vote exactly once in total, and a counted offset with c votes holds at least
    //        c - heavy light ones.  R2 is what decides reads over long homopolymers, whose repeated 7-mer
    //        sprays hundreds of votes over neighbouring offsets.
    int R = a.vub - c[0], top = c[0], R2 = a.hh + a.lp - max(0, c[0] - a.hh);
#pragma unroll
    for (int j = 1; j < NG; ++j) {
        int i0, i1, step = 1;
        if (j == 1) {
            i0 = nk - 1;
            i1 = nk - 1 - lim;
            step = -1;
        } else {
            // j = 2: middle; 3, 4: quarters; 5, 6: eighths next to the ends
            const int num = j == 2 ? 4 : j == 3 ? 2 : j == 4 ? 6 : j == 5 ? 1 : 7;
            i0 = (nk * num) >> 3;
            i1 = min(nk, i0 + lim);
        }
        g[j] = kNoCand;
        c[j] = 0;
        if (top > min(R, R2)) continue;        // already decided
        const int gj = unique_hit_offset(tab, head, rpk, i0, i1, step);
        bool dup = gj == kNoCand;
#pragma unroll
        for (int k = 0; k < NG; ++k)
            if (k < j && g[k] == gj) dup = true;
        if (dup) continue;
        g[j] = gj;
        c[j] = count_offset_call(a.rpk_off, a.hpk_off, nk, a.nk_hap, gj);
        R -= c[j];
        R2 -= max(0, c[j] - a.hh);
        top = max(top, c[j]);
    }
    if (!(top > min(R, R2)) || top == 0) return 0;
    int nt = 0;
    res[0] = res[1] = res[2] = (u32)kNoCand;
#pragma unroll
    for (int j = 0; j < NG; ++j)
        if (g[j] != kNoCand && c[j] == top) {
            if (nt < 3) res[nt] = (u32)(g[j] + 1);
            ++nt;
        }
    if (nt > 3) return 0;   // more tied maxima than the result holds: exact path
    // (the order of the tied offsets is irrelevant: the score is a min over the set, calign.pyx:239-247)
    return 1;
}

// kModes = false is the default instance (no flank score, no HLA clipping): the mode logic compiles away.
template <bool kModes, int kMinBlocks>
__global__ void __launch_bounds__(kAnchorThreads, kMinBlocks) k_anchor(DevBatch b, AnchorPlan plan, Queue q, ScoreParams sp_in,
                                                           Counters* ctr) {
    ScoreParams sp = sp_in;
    if (!kModes) sp.flank = sp.hla = 0;
    extern __shared__ __align__(16) uint8_t smem[];
    // areas at host-planned byte offsets (integer offsets keep every access a plain 32-bit shared
    // address; generic-pointer arithmetic costs an S2R/LEA sequence per use)
    u32* s_tab = (u32*)smem;
    u32* s_cnt = (u32*)(smem + plan.o_cnt);
    u32* s_fblist = (u32*)(smem + plan.o_fb);                 // pairs for the exact vote array
    u32* s_ulist = (u32*)(smem + plan.o_ul);                  // pairs the first guess left open
    u32* s_vlist = (u32*)(smem + plan.o_vl);                  // per pair: up to three tied-maximum offsets
    u32* s_rpk = (u32*)(smem + plan.o_rpk);                   // 2-bit packed reads
    u32* s_hpk = (u32*)(smem + plan.o_hpk);                   // 2-bit packed haplotypes (padded both sides)
    uint16_t* s_next = (uint16_t*)(smem + plan.o_next);
    u32* s_heavy = (u32*)(smem + plan.o_mult);                // bit per 7-mer key: repeats inside a haplotype of the sub-group
    uint16_t* s_heads = (uint16_t*)(smem + plan.o_heads);
    SlotInfo* s_slot = (SlotInfo*)(smem + plan.o_slot);
    int32_t* s_hmeta = (int32_t*)(smem + plan.o_hmeta);      // per hap: len, next offset, packed offset
    __shared__ int s_nid, s_nfb, s_nul, s_mmax, s_scan[kAnchorThreads / 32];

    const int tid = threadIdx.x, nthr = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
    unsigned long long st_pairs = 0, st_scored = 0, st_dp = 0, st_cells = 0, t_tile0 = 0;
    int w_prev = 0;
    __shared__ int s_tile;

    // Tiles are handed out through a counter (q.count[1]), not by block index: when this kernel shares the
    // GPU with another chunk's kernels not all of its CTAs are resident at once, and a static split would
    // leave the late CTAs' share for the end.
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(q.count + 1, 1);
        __syncthreads();
        const int ti = s_tile;
        if (tid == 0 && ctr) {   // longest tile so far, in microseconds (diagnostic: PlbRunStats.n_anchor_verify)
            unsigned long long now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t_tile0) atomicMax(&ctr->n_verify, (((now - t_tile0) / 1000ull) << 32) | (unsigned)w_prev);
            t_tile0 = now;
        }
        if (ti >= plan.n_tiles) break;
        const Tile tile = plan.tiles[ti];
        const int w = tile.w;
        w_prev = w;
        const int nh = tile.h1 - tile.h0;
        const int ns = (int)(tile.s1 - tile.s0);
        __syncthreads();
        if (tid == 0) {
            int noff = 0, poff = 0;
            for (int g = 0; g < nh; ++g) {
                const int h = tile.h0 + g;
                const int len = (int)(b.hap_seq_off[h + 1] - b.hap_seq_off[h]);
                s_hmeta[3 * g + 0] = len;
                s_hmeta[3 * g + 1] = noff;
                s_hmeta[3 * g + 2] = poff + kPackPadWords;   // word index of base 0
                noff += (len + 2) & ~1;
                poff += ((len + 15) >> 4) + 2 * kPackPadWords;
            }
            s_nfb = 0;
            s_nul = 0;
        }
        for (int i = tid; i < kRankWords; i += nthr) ((uint2*)s_tab)[i] = make_uint2(0u, 0u);
        // slot metadata + skip rule (chaplotype.pyx:343-361)
        for (int s = tid; s < ns; s += nthr) {
            const int64_t gs = tile.s0 + s;
            const int r = b.slot_read[gs];
            const int wi = b.slot_wi[gs];
            const int t = (int)(gs - b.wi_slot_off[wi]);
            SlotInfo si;
            si.read = r;
            si.len = (int)(b.read_seq_off[r + 1] - b.read_seq_off[r]);
            si.pos = b.read_pos[r];
            si.flags = 0;
            if (t < b.wi_n_good[wi] + b.wi_n_bad[wi]) {
                const int ov = read_overlap(b.win_start[w], b.win_end[w], si.pos, b.read_end[r]);
                if (b.read_qcfail[r] || ov < kKmer) si.flags = 1;
            }
            si.vub = 0;
            si.lph = 0;
            si.poff = 0;
            si.T = (int)(b.wi_slot_off[wi + 1] - b.wi_slot_off[wi]);
            si.pair0 = b.ll_off[wi] + t;
            s_slot[s] = si;
        }
        __syncthreads();
        if (tid < 32) {  // offsets of the packed read rows: warp scan
            int cp = 0;
            for (int s0 = 0; s0 < ns; s0 += 32) {
                const int s = s0 + tid;
                int np = 0;
                if (s < ns && s_slot[s].len > kKmer && !(s_slot[s].flags & 1))
                    np = ((s_slot[s].len + 15) >> 4) + kPackPadWords;
                int ip = np;
                for (int o = 1; o < 32; o <<= 1) {
                    const int vp = __shfl_up_sync(0xFFFFFFFFu, ip, o);
                    if (tid >= o) ip += vp;
                }
                if (s < ns) s_slot[s].poff = cp + ip - np;
                cp += __shfl_sync(0xFFFFFFFFu, ip, 31);
            }
        }
        __syncthreads();
        // ---- the only pass over the raw bases: pack reads and haplotypes 2 bits per base (the
        //      reference's hash digit, calign.pyx:69-74).  7-mer keys, read ids and vote counts all
        //      derive from these words. ----
        // (the loads of a pass are what this phase waits for, so every pass keeps as many of them in flight as it can: two
        // reads per 16 lanes at a time, and the haplotypes of the group as ONE index space - one haplotype after the other
        // left 22 of 256 threads busy per round trip to HBM)
        for (int s = tid >> 4; s < ns; s += nthr >> 3) {   // 16 lanes per read: a 150 bp read is 13 words
            const int s2 = s + (nthr >> 4);
            const SlotInfo sa = s_slot[s];
            const SlotInfo sb = s_slot[s2 < ns ? s2 : s];
            const bool va = !((sa.flags & 1) || sa.len <= kKmer);
            const bool vb = s2 < ns && !((sb.flags & 1) || sb.len <= kKmer);
            const uint8_t* ra = b.read_seq + b.read_seq_off[sa.read];
            const uint8_t* rb = b.read_seq + b.read_seq_off[sb.read];
            const int nwa = va ? ((sa.len + 15) >> 4) + kPackPadWords : 0;
            const int nwb = vb ? ((sb.len + 15) >> 4) + kPackPadWords : 0;
            for (int wd = tid & 15; wd < max(nwa, nwb); wd += 16) {
                const u32 xa = wd < nwa ? pack16_codes(ra, 16 * wd, sa.len) : 0u;
                const u32 xb = wd < nwb ? pack16_codes(rb, 16 * wd, sb.len) : 0u;
                if (wd < nwa) s_rpk[sa.poff + wd] = xa;
                if (wd < nwb) s_rpk[sb.poff + wd] = xb;
            }
        }
        {
            const int last = nh - 1;
            const int total = s_hmeta[3 * last + 2] + ((s_hmeta[3 * last] + 15) >> 4) + kPackPadWords;   // words of the group
            for (int k = tid; k < total; k += nthr) {
                int g = 0;
                while (g < last && k >= s_hmeta[3 * (g + 1) + 2] - kPackPadWords) ++g;
                const int len = s_hmeta[3 * g];
                const int wd = k - (s_hmeta[3 * g + 2] - kPackPadWords);
                s_hpk[k] = pack16_codes(b.hap_seq + b.hap_seq_off[tile.h0 + g], 16 * (wd - kPackPadWords), len);
            }
        }
        __syncthreads();
        // ---- union table: mark the key of every indexed haplotype position
        //      (calign.pyx:109: positions 0 .. len-8) ----
        for (int g = 0; g < nh; ++g) {
            const int nkh = s_hmeta[3 * g] - kKmer;
            const u32* hpk = s_hpk + s_hmeta[3 * g + 2];
            for (int i = tid; i < nkh; i += nthr) {
                const u32 key = key_at(hpk, i);
                atomicOr(&s_tab[2 * (key >> 5)], 1u << (key & 31u));
            }
        }
        __syncthreads();
        // ---- rank directory: exclusive prefix of the popcounts (block-wide scan over 512 words) ----
        {
            const int per = (kRankWords + nthr - 1) / nthr;
            const int lo = tid * per, hi = min(kRankWords, lo + per);
            int cnt = 0;
            for (int i = lo; i < hi; ++i) cnt += __popc(s_tab[2 * i]);
            int incl = cnt;
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if (lane >= o) incl += v;
            }
            if (lane == 31) s_scan[warp] = incl;
            __syncthreads();
            int base = 0;
            for (int k = 0; k < warp; ++k) base += s_scan[k];
            if (tid == nthr - 1) s_nid = base + incl;
            int run = base + incl - cnt;
            for (int i = lo; i < hi; ++i) {
                s_tab[2 * i + 1] = (u32)run;
                run += __popc(s_tab[2 * i]);
            }
        }
        __syncthreads();
        const int U = s_nid;  // number of distinct 7-mers in this haplotype group
        const int general = b.win_flags[w] & 1;
        const int hstride = (U + 2) & ~1;   // even: rows stay 4-byte aligned for the 32-bit CAS on u16 pairs
        const int sub_max = max(1, min(nh, plan.heads_halfs / hstride));
        for (int g0 = 0; g0 < nh; g0 += sub_max) {
            const int g1 = min(nh, g0 + sub_max);
            __syncthreads();
            // ---- chains of this haplotype sub-group (calign.pyx:94-124) ----
            for (int i = tid; i < (g1 - g0) * hstride; i += nthr) s_heads[i] = 0;
            for (int i = tid; i < kRankWords; i += nthr) s_heavy[i] = 0u;
            if (tid == 0) s_mmax = 1;
            __syncthreads();
            for (int g = g0; g < g1; ++g) {
                const int len = s_hmeta[3 * g];
                const u32* hpk = s_hpk + s_hmeta[3 * g + 2];
                uint16_t* nxt = s_next + s_hmeta[3 * g + 1];
                uint16_t* head = s_heads + (g - g0) * hstride;
                for (int i = tid; i < len - kKmer; i += nthr) {
                    const u32 key = key_at(hpk, i);
                    const u32 id = tab_lookup(s_tab, key);
                    unsigned short cur = head[id];
                    while (true) {  // push position i (stored as i+1); order inside a chain is irrelevant.
                        // bit 15 of the head marks chains with more than one element.
                        nxt[i + 1] = cur & 0x7FFFu;
                        const unsigned short nv = (unsigned short)((i + 1) | (cur ? 0x8000u : 0u));
                        const unsigned short old = cas_u16(head, (int)id, cur, nv);
                        if (old == cur) break;
                        cur = old;
                    }
                    // a 7-mer that repeats inside a haplotype of the sub-group is "heavy": it can vote several times
                    if (cur) atomicOr(&s_heavy[key >> 5], 1u << (key & 31u));
                }
            }
            __syncthreads();
            // largest multiplicity of a 7-mer inside one haplotype of the sub-group (chains with bit 15 only)
            for (int i = tid; i < (g1 - g0) * hstride; i += nthr) {
                const u32 hd = s_heads[i];
                if (hd & 0x8000u) {
                    const int g = g0 + i / hstride;
                    const uint16_t* nxt = s_next + s_hmeta[3 * g + 1];
                    int len = 0;
                    for (u32 p1 = hd & 0x7FFFu; p1; p1 = nxt[p1]) ++len;
                    atomicMax(&s_mmax, len);
                }
            }
            __syncthreads();
            // Per read: how many of its 7-mers (0..len-8, calign.pyx:155-165) occur in the group at all and how many are
            // heavy - two bitmap tests per 7-mer, five consecutive 7-mers per lane out of one 64-bit window of the packed
            // read.  light = present - heavy votes at most once on a haplotype; a heavy one votes at most s_mmax times:
            // V_ub = light + heavy * s_mmax bounds the votes the read can cast on any haplotype of the sub-group.
            const int mmax = s_mmax;
            for (int s = warp; s < ns; s += nwarp) {
                const SlotInfo si = s_slot[s];
                const int nk = si.len - kKmer;
                if ((si.flags & 1) || nk <= 0) continue;
                int np = 0, nhv = 0;
                const u32* rpk = s_rpk + si.poff;
                for (int base = 0; base < nk; base += 160) {
                    const int p0 = base + 5 * lane;
                    if (p0 < nk) {
                        const unsigned long long x =
                            (((unsigned long long)rpk[(p0 >> 4) + 1] << 32) | rpk[p0 >> 4]) >> (2 * (p0 & 15));
#pragma unroll
                        for (int t = 0; t < 5; ++t) {
                            if (p0 + t < nk) {
                                const u32 key = (u32)(x >> (2 * t)) & 0x3FFFu;
                                np += (int)((s_tab[2 * (key >> 5)] >> (key & 31u)) & 1u);
                                nhv += (int)((s_heavy[key >> 5] >> (key & 31u)) & 1u);
                            }
                        }
                    }
                }
                for (int o = 16; o > 0; o >>= 1) {
                    np += __shfl_xor_sync(0xFFFFFFFFu, np, o);
                    nhv += __shfl_xor_sync(0xFFFFFFFFu, nhv, o);
                }
                if (lane == 0) {
                    const int light = np - nhv;
                    s_slot[s].vub = light + nhv * mmax;
                    s_slot[s].lph = light | (nhv << 16);
                }
            }
            __syncthreads();
            // ---- per (slot, haplotype) pair: light decision, else exact vote array ----
            const int npairs = ns * (g1 - g0);
            const int hloc0 = tile.h0 - b.win_hap_off[w];   // index of the tile's first haplotype in its window
            const int hap_start_w = b.hap_start[w], win_start_w = b.win_start[w];
            // the flank score needs the scalar path for every alignment; HLA mode only for the pairs it clips
            const float inv_ns = 1.0f / (float)ns;   // p < 4096, ns <= 256: the float quotient is exact after truncation
            auto pair_id = [&](int p, int& s, int& g, int64_t& gs, int64_t& pair) {
                const int qd = (int)(((float)p + 0.5f) * inv_ns);
                s = p - qd * ns;
                g = g0 + qd;
                gs = tile.s0 + s;
                pair = s_slot[s].pair0 + (int64_t)(hloc0 + g) * s_slot[s].T;
            };
            for (int p = tid; p < npairs; p += nthr) {
                int s, g;
                int64_t gs, pair;
                pair_id(p, s, g, gs, pair);
                const SlotInfo si = s_slot[s];
                const PairClip pc = pair_clip(sp, si.pos, si.len, hap_start_w, win_start_w, s_hmeta[3 * g]);
                ++st_pairs;
                // a scored read must fit the haplotype (the reference would read past it, calign.pyx:256-259); the host
                // entry points refuse such batches, device-resident callers get the error flag and the sentinel score
                const bool too_long = !(si.flags & 1) && pc.L >= kKmer && pc.L + 15 > s_hmeta[3 * g];
                if (too_long) {
                    if (ctr) atomicOr(&ctr->err, kErrReadTooLong);
                    b.cand0[pair] = -1;
                    b.cand1[pair] = -1;
                    b.score[pair] = kScoreNone;
                    s_vlist[3 * p] = (u32)kPairSkip;
                    continue;
                }
                if ((si.flags & 1) || pc.L < kKmer) {  // LL forced to 0, or calign.pyx:182-183 (score 0)
                    b.cand0[pair] = -1;
                    b.cand1[pair] = -1;
                    b.score[pair] = (si.flags & 1) ? -1 : 0;
                    if (!(si.flags & 1)) {
                        ++st_scored;
                        st_cells += 16ull * (unsigned)(pc.L > 0 ? pc.L : 0);
                    }
                    s_vlist[3 * p] = (u32)kPairSkip;
                    continue;
                }
                ++st_scored;
                st_cells += 16ull * (unsigned)pc.L;
                LightArgs la;
                la.head_off = plan.o_heads + 2u * (u32)((g - g0) * hstride);
                la.rpk_off = plan.o_rpk + 4u * (u32)si.poff;
                la.hpk_off = plan.o_hpk + 4u * (u32)s_hmeta[3 * g + 2];
                la.res_off = plan.o_vl + 12u * (u32)p;
                la.nk_read = pc.L - kKmer;   // HLA mode: 7-mers 0..L'-8 of the UNCLIPPED read vote
                la.nk_hap = s_hmeta[3 * g] - kKmer;
                la.vub = si.vub;
                la.lp = si.lph & 0xFFFF;
                la.hh = si.lph >> 16;
                if (!light_first(la)) s_ulist[atomicAdd(&s_nul, 1)] = (u32)p;
            }
            __syncthreads();
            // ---- step two: the pairs left open, compacted (their order in the list does not matter) ----
            const int nul = s_nul;
            if (tid == 0 && ctr) atomicAdd(&ctr->n_heavy, (unsigned long long)nul);   // diagnostic: pairs needing step two
            for (int k = tid; k < nul; k += nthr) {
                const int p = (int)s_ulist[k];
                int s, g;
                int64_t gs, pair;
                pair_id(p, s, g, gs, pair);
                const SlotInfo si = s_slot[s];
                const PairClip pc = pair_clip(sp, si.pos, si.len, hap_start_w, win_start_w, s_hmeta[3 * g]);
                LightArgs la;
                la.head_off = plan.o_heads + 2u * (u32)((g - g0) * hstride);
                la.rpk_off = plan.o_rpk + 4u * (u32)si.poff;
                la.hpk_off = plan.o_hpk + 4u * (u32)s_hmeta[3 * g + 2];
                la.res_off = plan.o_vl + 12u * (u32)p;
                la.nk_read = pc.L - kKmer;
                la.nk_hap = s_hmeta[3 * g] - kKmer;
                la.vub = si.vub;
                la.lp = si.lph & 0xFFFF;
                la.hh = si.lph >> 16;
                if (!light_rest(la)) {
                    s_vlist[3 * p] = (u32)kPairUndecided;
                    s_fblist[atomicAdd(&s_nfb, 1)] = (u32)p;
                }
            }
            __syncthreads();
            // ---- emit the band starts of the decided pairs ----
            for (int p = tid; p < npairs; p += nthr) {
                const int c0 = (int)s_vlist[3 * p];
                if (c0 == kPairSkip || c0 == kPairUndecided) continue;
                int s, g;
                int64_t gs, pair;
                pair_id(p, s, g, gs, pair);
                const SlotInfo si = s_slot[s];
                const int hap_len = s_hmeta[3 * g], h = tile.h0 + g;
                const PairClip pc = pair_clip(sp, si.pos, si.len, hap_start_w, win_start_w, hap_len);
                const int L = pc.L, roff = pc.off1;
                Emitter<kModes> em;
                em.c0 = em.c1 = -1;
                em.sc = kScoreNone;
                em.n_dp = 0;
                int idx0 = si.pos + roff - hap_start_w;  // fallback position (calign.pyx:252-256)
                const int lim = hap_len - L - 15;
                if (lim < idx0) idx0 = lim;
                const bool slow = general || sp.flank || (sp.hla && (roff != 0 || L != si.len)) || L < kMinFastLen ||
                                  L > kMaxFastLen;
                bool any_accepted = false;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    const int cj = (int)s_vlist[3 * p + j];
                    if (cj == kNoCand) continue;
                    const int idx = cj - 1;
                    if (idx + L + 15 < hap_len) {  // calign.pyx:228
                        any_accepted = true;
                        em.emit(b, q, sp, slow, pair, h, gs, si.read, roff, L, idx > 8 ? idx - 8 : 0);
                    }
                }
                // with no accepted candidate bestMappingPosition stays -1, so a fallback index of
                // exactly -1 is skipped and the sentinel 1000000 is returned (calign.pyx:258, 272)
                if (any_accepted || idx0 != -1)
                    em.emit(b, q, sp, slow, pair, h, gs, si.read, roff, L, idx0 > 8 ? idx0 - 8 : 0);
                st_dp += em.n_dp;
                b.cand0[pair] = em.c0;
                b.cand1[pair] = em.c1;
                b.score[pair] = em.sc;
            }
            __syncthreads();
            // ---- undecided pairs: exact vote array per warp (calign.pyx:206-247) ----
            const int nfb = s_nfb;
            if (tid == 0 && ctr) {
                atomicAdd(&ctr->n_exact, (unsigned long long)nfb);
            }
            const int n_fbw = min(nwarp, plan.n_cnt);  // warps that own a counter array
            for (int f = warp; f < nfb && warp < n_fbw; f += n_fbw) {
                const int p = (int)s_fblist[f];
                const int qd = (int)(((float)p + 0.5f) * inv_ns);
                const int s = p - qd * ns, g = g0 + qd;
                const SlotInfo si = s_slot[s];
                const int h = tile.h0 + g;
                const int64_t gs = tile.s0 + s;
                const int64_t pair = si.pair0 + (int64_t)(h - b.win_hap_off[w]) * si.T;
                const int hap_len = s_hmeta[3 * g];
                const PairClip pc = pair_clip(sp, si.pos, si.len, hap_start_w, win_start_w, hap_len);
                const int L = pc.L, roff = pc.off1, nk = L - kKmer;
                const uint16_t* nxt = s_next + s_hmeta[3 * g + 1];
                const uint16_t* head = s_heads + (g - g0) * hstride;
                const u32* rpk = s_rpk + si.poff;
                u32* cw = s_cnt + (size_t)warp * plan.cnt_words;
                const int C = hap_len + L;
                const int words = (C + 1) >> 1;
                for (int k = lane; k < words; k += 32) cw[k] = 0;
                __syncwarp();
                for (int i = lane; i < nk; i += 32) {
                    const u32 id = tab_lookup(s_tab, key_at(rpk, i));
                    if (!id) continue;
                    u32 p1 = head[id] & 0x7FFFu;
                    while (p1) {
                        const int o = (int)p1 - i - 1 + L;  // index pos + readLen, calign.pyx:213-215
                        atomicAdd(&cw[o >> 1], 1u << (16 * (o & 1)));
                        p1 = nxt[p1];
                    }
                }
                __syncwarp();
                u32 m = 0;
                for (int k = lane; k < words; k += 32) {
                    const u32 v = cw[k];
                    m = max(m, max(v & 0xFFFFu, v >> 16));
                }
                for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
                Emitter<kModes> em;
                em.c0 = em.c1 = -1;
                em.sc = kScoreNone;
                em.n_dp = 0;
                const bool slow = general || sp.flank || (sp.hla && (roff != 0 || L != si.len)) || L < kMinFastLen ||
                                  L > kMaxFastLen;
                bool any_accepted = false;
                for (int base = 0; base < C; base += 32) {  // ascending offsets, calign.pyx:223
                    const int o = base + lane;
                    u32 c = 0;
                    if (o < C) c = (cw[o >> 1] >> (16 * (o & 1))) & 0xFFFFu;
                    const int idx = o - L;
                    const bool hit = (o < C) && m > 0 && c == m && (idx + L + 15 < hap_len);
                    unsigned bal = __ballot_sync(0xFFFFFFFFu, hit);
                    if (lane == 0) {
                        while (bal) {
                            const int k = __ffs(bal) - 1;
                            bal &= bal - 1;
                            const int ix = base + k - L;
                            any_accepted = true;
                            em.emit(b, q, sp, slow, pair, h, gs, si.read, roff, L, ix > 8 ? ix - 8 : 0);
                        }
                    }
                }
                if (lane == 0) {
                    int idx0 = si.pos + roff - hap_start_w;
                    const int lim = hap_len - L - 15;
                    if (lim < idx0) idx0 = lim;
                    if (any_accepted || idx0 != -1)
                        em.emit(b, q, sp, slow, pair, h, gs, si.read, roff, L, idx0 > 8 ? idx0 - 8 : 0);
                    st_dp += em.n_dp;
                    b.cand0[pair] = em.c0;
                    b.cand1[pair] = em.c1;
                    b.score[pair] = em.sc;
                }
                __syncwarp();
            }
            __syncthreads();
            if (tid == 0) {
                s_nfb = 0;
                s_nul = 0;
            }
        }
    }
    if (ctr) {  // statistics
        for (int o = 16; o > 0; o >>= 1) {
            st_pairs += __shfl_down_sync(0xFFFFFFFFu, st_pairs, o);
            st_scored += __shfl_down_sync(0xFFFFFFFFu, st_scored, o);
            st_dp += __shfl_down_sync(0xFFFFFFFFu, st_dp, o);
            st_cells += __shfl_down_sync(0xFFFFFFFFu, st_cells, o);
        }
        if (lane == 0) {
            atomicAdd(&ctr->n_pairs, st_pairs);
            atomicAdd(&ctr->n_scored, st_scored);
            atomicAdd(&ctr->n_dp, st_dp);
            atomicAdd(&ctr->cells, st_cells);
        }
    }
}

// One band alignment by 16 lanes: lane d owns diagonal d and the lanes sweep anti-diagonals in lock step,
// exchanging neighbours' states by warp shuffles (the reference's SSE2 scheme, src/c/align.c:199-515, with
// one lane per diagonal instead of two interleaved 8-lane vectors).  Cell (x, y) has x - y = d and is
// computed at step s = x + y, so a lane is active on every other step:
//   M(x,y) <- own diagonal, two steps ago        I(x,y) <- lane d+1, previous step        D(x,y) <- lane d-1
// Latency per alignment is ~2L steps instead of the 16L cells of the one-thread scalar form, which is what
// the queue needs: few alignments, nothing to hide their latency behind.  Same recurrence as
// band_dp_general (byte-exact compares, any read length >= 1).  All 32 lanes of the warp must call it;
// `sub` selects the half-warp's alignment (lanes 0-15 / 16-31), inactive halves pass L = 0.
__device__ __forceinline__ int band_dp_wave16(const uint8_t* __restrict__ hap, const uint8_t* __restrict__ open,
                                              const uint8_t* __restrict__ read, const uint8_t* __restrict__ qual, int L,
                                              int Lmax, int ext, int nuc) {
    const unsigned FULL = 0xFFFFFFFFu;
    const int d = threadIdx.x & 15;
    int M = kScoreBig, I = kScoreBig, D = kScoreBig;   // this diagonal's most recent cell
    const int n_steps = 2 * (Lmax - 1) + 16;           // uniform over the warp
    // bytes of the lane's next cell, loaded two steps ahead (its cells are (d,0), (d+1,1), ...)
    int hb = 0, go = 0, rb = 0, q = 0;
    if (L > 0) {
        hb = hap[d];
        go = open[d];
        rb = read[0];
        q = qual[0];
    }
    for (int s = 0; s < n_steps; ++s) {
        // neighbours' states as of the previous step (lane d+1: cell (x, y-1); lane d-1: cell (x-1, y))
        const int upI = __shfl_down_sync(FULL, I, 1, 16), upM = __shfl_down_sync(FULL, M, 1, 16);
        const int lfD = __shfl_up_sync(FULL, D, 1, 16);
        const int mi = M < I ? M : I;
        const int lfMI = __shfl_up_sync(FULL, mi, 1, 16);
        const int y2 = s - d;
        if (y2 < 0 || (y2 & 1)) continue;
        const int y = y2 >> 1, x = y + d;
        if (y >= L) continue;
        const int chb = hb, cgo = go, crb = rb, cq = q;
        if (y + 1 < L) {   // prefetch the next cell of this diagonal
            hb = hap[x + 1];
            go = open[x + 1];
            rb = read[y + 1];
            q = qual[y + 1];
        }
        const int sub = (chb == 'N' || chb == crb) ? 0 : cq;
        int diag = mi < D ? mi : D;                     // B(x-1, y-1): own state from two steps ago
        if (y == 0) diag = 0;
        const int m = diag + sub;
        int ins;
        if (y == 0) {
            ins = (x & 1) ? kScoreBig : cgo + nuc;
        } else if (d < 15) {
            const int a = upI + ext, c = upM + cgo;
            ins = (a < c ? a : c) + nuc;
        } else {
            ins = kScoreBig;
        }
        int del;
        if (d >= 1) {
            const int a = lfD + ext, c = lfMI + cgo;
            del = a < c ? a : c;
        } else {
            del = kScoreBig;
        }
        M = m < kScoreBig ? m : kScoreBig;
        I = ins < kScoreBig ? ins : kScoreBig;
        D = del < kScoreBig ? del : kScoreBig;
    }
    int best = M < I ? M : I;
    best = best < D ? best : D;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        const int v = __shfl_xor_sync(FULL, best, o, 16);
        best = v < best ? v : best;
    }
    return best;
}

// ---------------------------------------------------------------------------------------------
// k_general: queued alignments.  Default mode: 16 lanes per alignment (band_dp_wave16); run-time modes
// (flank score / HLA clipping, where EVERY alignment is queued): one thread per alignment.
// ---------------------------------------------------------------------------------------------
template <bool kModes>
__global__ void __launch_bounds__(128) k_general(DevBatch b, Queue q, ScoreParams sp) {
    int n = *q.count;
    if (n > q.cap) n = q.cap;
    if (kModes) {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
            const QueueEntry qe = q.e[i];
            const int r = b.slot_read[qe.slot];
            int L = (int)(b.read_seq_off[r + 1] - b.read_seq_off[r]), roff = 0;
            if (sp.hla) {
                roff = (int)(qe.clip & 0xFFFFu);
                L = (int)(qe.clip >> 16);
            }
            const int v = general_dp_now<true>(b, qe.hap, r, qe.start, roff, L, sp);
            atomicMin(&b.score[qe.pair], v);
        }
        return;
    }
    const int hw = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;     // global half-warp index
    const int n_hw = (gridDim.x * blockDim.x) >> 4;
    const int n_round = (n + n_hw - 1) / n_hw;
    for (int k = 0; k < n_round; ++k) {                               // uniform trip count per warp
        const int i = k * n_hw + hw;
        int L = 0;
        const uint8_t *hap = nullptr, *go = nullptr, *rs = nullptr, *rq = nullptr;
        int64_t pair = 0;
        if (i < n) {
            const QueueEntry qe = q.e[i];
            const int r = b.slot_read[qe.slot];
            L = (int)(b.read_seq_off[r + 1] - b.read_seq_off[r]);
            hap = b.hap_seq + b.hap_seq_off[qe.hap] + qe.start;
            go = b.gap_open + b.hap_seq_off[qe.hap] + qe.hap + qe.start;
            rs = b.read_seq + b.read_seq_off[r];
            rq = b.read_qual + b.read_seq_off[r];
            pair = qe.pair;
        }
        const int Lo = __shfl_xor_sync(0xFFFFFFFFu, L, 16);
        const int Lmax = L > Lo ? L : Lo;
        if (Lmax == 0) continue;                                      // both halves idle (warp-uniform)
        const int v = band_dp_wave16(hap, go, rs, rq, L, Lmax, sp.ext, sp.nuc);
        if (i < n && L > 0 && (threadIdx.x & 15) == 0) atomicMin(&b.score[pair], v);
    }
}

// ---------------------------------------------------------------------------------------------
// k_dp
// ---------------------------------------------------------------------------------------------
// ---- TMA (bulk async copy) + mbarrier wrappers: cp.async.bulk global -> shared, completion counted
//      in bytes on an mbarrier.  SASS: UBLKCP / SYNCS. ----
__device__ __forceinline__ u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, u32 count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, u32 bytes) {
    asm volatile("{\n.reg .b64 st;\nmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_addr(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, u32 parity) {
    asm volatile(
        "{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(
            smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* smem_dst, const void* gmem_src, u32 bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(smem_dst)),
                 "l"(__cvta_generic_to_global(gmem_src)), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// words of one profile row in shared memory (same rule as the host planner)
__device__ __forceinline__ int prof_row_words_dev(int L) {
    int n = dp_steps(L) + 4;           // rows read by the kernel, +4 keeps rows 16-byte apart
    n = (n + 3) & ~3;
    if (!((n >> 2) & 1)) n += 4;       // odd multiple of 16 bytes: conflict-free LDS.128
    return n;
}
constexpr int kProfTabWords = 5 * 128;   // k_dp's profile table: 5 base codes x 128 qualities
constexpr int kTmaMaxLen = 192;   // rows up to this length are staged by TMA (raw bytes fit the row tail)
__device__ __forceinline__ int tma_raw_bytes(int L) { return (L + 30 + 15) & ~15; }   // per array, 16-byte multiple

struct DpPlan {
    const Tile* tiles;
    int32_t n_tiles;
    int32_t max_slots;
    int32_t max_group;
    int32_t prof_words;   // smem u32 words for all profile rows of a tile
    int32_t rec_count;    // smem HapRec entries for a haplotype group
    int32_t max_pairs;    // max slots*group per tile
    int32_t tma_max;      // reads up to this length are staged by TMA (<= kTmaMaxLen; PLB_DP_TMA_MAX, experiments)
};

struct DpSlot {
    int32_t read;
    int32_t len;
    int32_t poff;    // offset of the profile row (u32 words)
    int32_t flags;
    double ll_right; // log(1 - exp(mLTOT*mapq)), chaplotype.pyx:622
};

// k_read_check: one warp per read of the pool, ONCE per read (a read shared by several windows is scored in several
// tiles).  The packed int16 recurrences are exact while every path cost stays inside the reference's own range
// (pos_inf >> 2 = 15,872 phred, align.c:97; a score never exceeds the sum of the read's qualities plus gap costs that the
// margin below covers); a read whose qualities add up beyond that takes the 32-bit recurrence instead (flag bit 1).  A
// quality above 93 is an input error (the reference asserts it, htslibWrapper.pyx:518-519).
constexpr int kMaxPackedQualSum = 15871 - 256;
__global__ void __launch_bounds__(256) k_read_check(DevBatch b, int r0, int r1, Counters* __restrict__ ctr) {
    const int r = r0 + (int)((int64_t)blockIdx.x * blockDim.x + threadIdx.x);
    if (r >= r1) return;
    // one thread per read: aligned 32-bit loads, bytes outside the read masked off, four qualities summed / maximised per
    // instruction (the arrays are padded by 64 bytes, so the last word may be read whole)
    const int64_t o = b.read_seq_off[r];
    const int L = (int)(b.read_seq_off[r + 1] - o);
    const uintptr_t a0 = (uintptr_t)(b.read_qual + o);
    const u32* wp = (const u32*)(a0 & ~(uintptr_t)3);
    const int head = (int)(a0 & 3);                  // bytes of the first word that belong to the previous read
    const int nw = (head + L + 3) >> 2;
    u32 sum = 0, mx = 0;
    for (int k = 0; k < nw; ++k) {
        u32 v = __ldg(wp + k);
        if (k == 0) v &= 0xFFFFFFFFu << (8 * head);
        if (k == nw - 1) {
            const int tail = 4 * nw - (head + L);    // bytes of the last word past the read
            v &= 0xFFFFFFFFu >> (8 * tail);
        }
        sum += __vsadu4(v, 0u);
        mx = __vmaxu4(mx, v);
    }
    const u32 m2 = max(max(mx & 0xFFu, (mx >> 8) & 0xFFu), max((mx >> 16) & 0xFFu, mx >> 24));
    b.read_flags[r] = (L > 0 && sum > (u32)kMaxPackedQualSum) ? 2 : 0;
    if (m2 > 93u && ctr) atomicOr(&ctr->err, kErrQuality);
}

template <int NTHR>
__global__ void __launch_bounds__(NTHR, 3) k_dp(DevBatch b, DpPlan plan, ScoreParams sp, double* __restrict__ ll_out,
                                             int32_t* __restrict__ score_out, int* __restrict__ tile_counter,
                                             Counters* __restrict__ ctr) {
    extern __shared__ __align__(16) uint8_t smem[];
    u32* s_prof = (u32*)smem;
    HapRec* s_rec = (HapRec*)(s_prof + plan.prof_words);
    DpSlot* s_slot = (DpSlot*)(s_rec + plan.rec_count);
    int32_t* s_roff = (int32_t*)(s_slot + plan.max_slots);       // per hap: record offset
    int32_t* s_order = s_roff + plan.max_group;                  // slots by decreasing read length
    int32_t* s_best = s_order + plan.max_slots;                  // per pair: best score
    u32* s_task = (u32*)(s_best + plan.max_pairs);               // compacted tasks
    u32* s_ptab = s_task + 2 * plan.max_pairs;                   // profile word by (base code, quality): [5][128]
    uint8_t* s_code = (uint8_t*)(s_ptab + kProfTabWords);        // byte -> base code 0..3 (exact A/C/G/T) or 4
    __shared__ int s_ntask;
    __shared__ __align__(8) uint64_t s_bar;   // counts the bytes of the tile's TMA copies

    const int tid = threadIdx.x;
    if (tid == 0) mbar_init(&s_bar, NTHR);
    u32 bar_phase = 0;
    for (int c = tid; c < 256; c += NTHR) {
        const int fc = fast_code((uint8_t)c);
        s_code[c] = (uint8_t)(fc < 4 ? fc : 4);
    }
    int ptab_mode = -1;   // which variant (6-op / 8-op costs) the profile table currently holds
    __syncthreads();
    __shared__ int s_tile;
    while (true) {   // dynamic tile hand-out, see k_anchor
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(tile_counter, 1);
        __syncthreads();
        const int ti = s_tile;
        if (ti >= plan.n_tiles) break;
        const Tile tile = plan.tiles[ti];
        const int w = tile.w;
        const int nh = tile.h1 - tile.h0;
        const int ns = (int)(tile.s1 - tile.s0);
        if (tid == 0) {
            s_ntask = 0;
            int ro = 0;
            for (int g = 0; g < nh; ++g) {
                const int h = tile.h0 + g;
                s_roff[g] = ro;
                ro += (int)(b.hap_seq_off[h + 1] - b.hap_seq_off[h]) + kRecPad;
            }
        }
        int my_exact = 0;
        for (int s = tid; s < ns; s += NTHR) {
            const int64_t gs = tile.s0 + s;
            const int r = b.slot_read[gs];
            const int wi = b.slot_wi[gs];
            const int t = (int)(gs - b.wi_slot_off[wi]);
            DpSlot ds;
            ds.read = r;
            ds.len = (int)(b.read_seq_off[r + 1] - b.read_seq_off[r]);
            ds.flags = 0;
            if (t < b.wi_n_good[wi] + b.wi_n_bad[wi]) {
                const int ov = read_overlap(b.win_start[w], b.win_end[w], b.read_pos[r], b.read_end[r]);
                if (b.read_qcfail[r] || ov < kKmer) ds.flags = 1;
            }
            if (b.read_flags[r] & 2) {   // qualities beyond the int16 range (k_read_check): exact 32-bit recurrence
                ds.flags |= 2;
                my_exact = 1;
            }
            ds.flags |= (int)b.read_mapq[r] << 8;
            ds.ll_right = c_map_right[b.read_mapq[r]];
            ds.poff = 0;
            s_slot[s] = ds;
        }
        const int any_exact = __syncthreads_or(my_exact);   // block-uniform
        // Tasks are enumerated slot-major in order of decreasing read length, so that the 32 alignments
        // of a warp have (almost) the same number of steps and the longest ones start first.
        for (int s = tid; s < ns; s += NTHR) {
            const int L = s_slot[s].len;
            int rank = 0;
            for (int j = 0; j < ns; ++j) {
                const int Lj = s_slot[j].len;
                rank += (Lj > L) || (Lj == L && j < s);
            }
            s_order[rank] = s;
        }
        if (tid < 32) {  // exclusive scan of the profile-row sizes by one warp
            int carry = 0;
            for (int s0 = 0; s0 < ns; s0 += 32) {
                const int s = s0 + tid;
                int n = 0;
                if (s < ns) {
                    const int L = s_slot[s].len;
                    if (!(s_slot[s].flags & 1) && L >= kMinFastLen && L <= kMaxFastLen) n = prof_row_words_dev(L);
                }
                int incl = n;
                for (int o = 1; o < 32; o <<= 1) {
                    const int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                    if (tid >= o) incl += v;
                }
                if (s < ns) s_slot[s].poff = carry + incl - n;
                carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
            }
        }
        __syncthreads();
        const u32 wflags = b.win_flags[w];
        // with the flank score every alignment runs on the scalar path (k_general) and this kernel only turns
        // the scores into log-likelihoods; HLA mode sends only the pairs it clips there
        const int general = (wflags & 1) | sp.flank;
        const bool six = !(wflags & 2);   // no 'N' in the window's haplotypes: 6-op variant
        const bool five = six && !(wflags & 4);   // ... and every gap-open >= ext: 5-op variant (one VIMNMX3)
        const int K = 2 * sp.ext + sp.nuc;
        // Profile words come from a table indexed by (base code, quality): two LDS instead of a dozen ALU ops
        // per read base (the ALU pipe is this kernel's bottleneck, the LSU pipe is nearly idle).
        if (!general && ptab_mode != (int)six) {   // warp-uniform: wflags is per tile
            for (int e = tid; e < kProfTabWords; e += NTHR) {
                const int c = e >> 7, q = e & 127;
                s_ptab[e] = six ? make_profile6(c < 4 ? c : 5, q, K) : make_profile(c < 4 ? c : 5, (u32)q);
            }
            ptab_mode = (int)six;   // visible to the profile pass after the barriers below
        }
        // ---- TMA: one bulk copy per read for bases and one for qualities, straight into the tail of
        //      the read's profile row; every thread arrives on the mbarrier, copies add their bytes ----
        {
            bool issued = false;
            if (!general && tid < ns) {
                const DpSlot ds = s_slot[tid];
                if (!(ds.flags & 1) && ds.len >= kMinFastLen && ds.len <= plan.tma_max) {
                    const int64_t o = b.read_seq_off[ds.read];
                    const int64_t a0 = o & ~(int64_t)15;
                    const u32 nb = (u32)(((o + ds.len + 15) & ~(int64_t)15) - a0);   // <= tma_raw_bytes(len)
                    uint8_t* row_end = (uint8_t*)(s_prof + ds.poff + prof_row_words_dev(ds.len));
                    const int NB = tma_raw_bytes(ds.len);
                    fence_proxy_async();   // the row was read through the generic proxy by the previous tile
                    mbar_arrive_expect_tx(&s_bar, 2 * nb);
                    tma_load_1d(row_end - 2 * NB, b.read_seq + a0, nb, &s_bar);
                    tma_load_1d(row_end - NB, b.read_qual + a0, nb, &s_bar);
                    issued = true;
                }
            }
            if (!issued) mbar_arrive(&s_bar);
        }
        // haplotype records (while the read bytes are in flight)
        if (!general) {
            for (int g = 0; g < nh; ++g) {
                const int h = tile.h0 + g;
                const int len = (int)(b.hap_seq_off[h + 1] - b.hap_seq_off[h]);
                const uint8_t* hap = b.hap_seq + b.hap_seq_off[h];
                const uint8_t* go = b.gap_open + b.hap_seq_off[h] + h;
                HapRec* rec = s_rec + s_roff[g];
                for (int x = tid; x < len + kRecPad; x += NTHR) {
                    const uint8_t ha = x < len ? hap[x] : (uint8_t)'N', hb = x + 4 < len ? hap[x + 4] : (uint8_t)'N';
                    const u32 oa = x <= len ? go[x] : 0u;
                    const u32 ob = x + 4 <= len ? go[x + 4] : 0u;
                    const int ca = fast_code(ha), cb = fast_code(hb);
                    HapRec r;
                    if (six) {
                        r.gow = pack_s16x2((int)oa - sp.ext, (int)ob - sp.ext);
                        r.sel = make_sel6(ca < 4 ? ca : 0, cb < 4 ? cb : 0);
                    } else {
                        r.gow = oa | (ob << 16);
                        r.sel = make_sel(ca, cb);
                    }
                    rec[x] = r;
                }
            }
        }
        mbar_wait(&s_bar, bar_phase);
        bar_phase ^= 1;
        // profiles: one warp per read row, lanes along the read
        if (!general) {
            const int warp = tid >> 5, lane = tid & 31, nwarp = NTHR >> 5;
            for (int s = warp; s < ns; s += nwarp) {
                const DpSlot ds = s_slot[s];
                if ((ds.flags & 1) || ds.len < kMinFastLen || ds.len > kMaxFastLen) continue;
                const int64_t o = b.read_seq_off[ds.read];
                int n = dp_steps(ds.len) + 4;
                u32* row = s_prof + ds.poff;
                if (ds.len <= plan.tma_max) {
                    // raw bytes sit in the row's own tail: pull all of them into registers, then overwrite
                    const uint8_t* row_end = (const uint8_t*)(row + prof_row_words_dev(ds.len));
                    const int NB = tma_raw_bytes(ds.len);
                    const uint8_t* rs = row_end - 2 * NB + (int)(o & 15);
                    const uint8_t* rq = row_end - NB + (int)(o & 15);
                    uint8_t cb[6], qb[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int y = lane + 32 * k;
                        cb[k] = y < ds.len ? rs[y] : (uint8_t)0;
                        qb[k] = y < ds.len ? rq[y] : (uint8_t)0;
                    }
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int y = lane + 32 * k;
                        if (y < n) {
                            row[y] = y < ds.len ? s_ptab[((u32)s_code[cb[k]] << 7) | (qb[k] & 127u)] : 0u;
                        }
                    }
                    // reads of 177-192 bp need rows beyond the 192 handled above: the recurrence reads dp_steps(L) + 4 rows
                    // and the rows past the read MUST be zero (stale words there can wrap a masked lane negative, and a
                    // negative value survives the last-row mask: found by the 1974-window differential case)
                    for (int y = 192 + lane; y < n; y += 32) row[y] = 0u;
                    continue;
                }
                const uint8_t* rs = b.read_seq + o;
                const uint8_t* rq = b.read_qual + o;
                for (int y0 = 0; y0 < n; y0 += 192) {   // long reads: plain loads, 6 rows per lane
                    uint8_t cb[6], qb[6];
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int y = y0 + lane + 32 * k;
                        cb[k] = y < ds.len ? rs[y] : (uint8_t)0;
                        qb[k] = y < ds.len ? rq[y] : (uint8_t)0;
                    }
#pragma unroll
                    for (int k = 0; k < 6; ++k) {
                        const int y = y0 + lane + 32 * k;
                        if (y < n) {
                            row[y] = y < ds.len ? s_ptab[((u32)s_code[cb[k]] << 7) | (qb[k] & 127u)] : 0u;
                        }
                    }
                }
            }
        }
        // best scores start from what the general path produced; compact packed-path tasks
        const int npairs = ns * nh;
        for (int p0 = 0; p0 < npairs; p0 += NTHR) {
            const int p = p0 + tid;
            int c0 = -1, c1 = -1;
            if (p < npairs) {
                const int s = s_order[p / nh], g = p % nh;
                const int64_t gs = tile.s0 + s;
                const int wi = b.slot_wi[gs];
                const int64_t T = b.wi_slot_off[wi + 1] - b.wi_slot_off[wi];
                const int64_t pair =
                    b.ll_off[wi] + (int64_t)(tile.h0 + g - b.win_hap_off[w]) * T + (gs - b.wi_slot_off[wi]);
                s_best[p] = b.score[pair];
                c0 = b.cand0[pair];
                c1 = b.cand1[pair];
            }
            const int cnt = (c0 >= 0) + (c1 >= 0);
            // block-wide exclusive scan via warp ballots would need two levels; an atomic per
            // warp keeps it simple (order of tasks inside a tile does not affect the result)
            const unsigned lane = tid & 31;
            int incl = cnt;
            for (int o = 1; o < 32; o <<= 1) {
                int v = __shfl_up_sync(0xFFFFFFFFu, incl, o);
                if ((int)lane >= o) incl += v;
            }
            int base = 0;
            if (lane == 31) base = atomicAdd(&s_ntask, incl);
            base = __shfl_sync(0xFFFFFFFFu, base, 31);
            int pos = base + incl - cnt;
            if (c0 >= 0) s_task[pos++] = ((u32)p << 15) | (u32)c0;
            if (c1 >= 0) s_task[pos++] = ((u32)p << 15) | (u32)c1;
        }
        __syncthreads();
        const int ntask = s_ntask;
        for (int k = tid; k < ntask; k += NTHR) {
            const u32 tk = s_task[k];
            const int p = (int)(tk >> 15), start = (int)(tk & 0x7FFFu);
            const int s = s_order[p / nh], g = p % nh;
            const DpSlot ds = s_slot[s];
            if (ds.flags & 2) continue;   // qualities beyond the int16 range: second pass below
            const int v = five  ? band_dp_fast5(s_prof + ds.poff, s_rec + s_roff[g] + start, ds.len, sp.ext, sp.nuc)
                          : six ? band_dp_fast6(s_prof + ds.poff, s_rec + s_roff[g] + start, ds.len, sp.ext, sp.nuc)
                                : band_dp_fast(s_prof + ds.poff, s_rec + s_roff[g] + start, ds.len, sp.ext, sp.nuc);
            atomicMin(&s_best[p], v);
        }
        if (any_exact) {   // block-uniform: the exact 32-bit recurrence, out of line
            for (int k = tid; k < ntask; k += NTHR) {
                const u32 tk = s_task[k];
                const int p = (int)(tk >> 15), start = (int)(tk & 0x7FFFu);
                const int s = s_order[p / nh], g = p % nh;
                if (!(s_slot[s].flags & 2)) continue;
                const int v = general_dp_now<false>(b, tile.h0 + g, s_slot[s].read, start, 0, s_slot[s].len, sp);
                atomicMin(&s_best[p], v);
            }
        }
        __syncthreads();
        // score -> log-likelihood (chaplotype.pyx:675-676) and output
        for (int p = tid; p < npairs; p += NTHR) {
            const int s = s_order[p / nh], g = p % nh;
            const DpSlot ds = s_slot[s];
            const int64_t gs = tile.s0 + s;
            const int wi = b.slot_wi[gs];
            const int64_t T = b.wi_slot_off[wi + 1] - b.wi_slot_off[wi];
            const int64_t pair =
                b.ll_off[wi] + (int64_t)(tile.h0 + g - b.win_hap_off[w]) * T + (gs - b.wi_slot_off[wi]);
            const int sc = s_best[p];
            double ll;
            if (ds.flags & 1) {
                ll = 0.0;
            } else if (sp.hla) {
                // chaplotype.pyx:631-634, 664-676: the cap is the log-probability of a wrong mapping and
                // scores above 100 are flattened: mLTOT * (99 + (score - 99)^0.5 / 0.5)
                const double cap = kMLTOT * (double)((ds.flags >> 8) & 0xFF);
                const double v = sc > 100 ? kMLTOT * (99.0 + sqrt((double)sc - 99.0) / 0.5)
                                          : __dadd_rn(__dmul_rn(kMLTOT, (double)sc), ds.ll_right);
                ll = v > cap ? v : cap;
            } else {
                const double v = __dadd_rn(__dmul_rn(kMLTOT, (double)sc), ds.ll_right);   // not fused, see c_map_right
                ll = v > -300.0 ? v : -300.0;
            }
            if (ll_out) ll_out[pair] = ll;
            if (score_out) score_out[pair] = sc;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_genotype: one block per (window, individual)
// ---------------------------------------------------------------------------------------------
struct PopOut {
    int32_t max_haps;
    double* gl;
    double* gl_log_max;
    double* gof;
    double* hap_like;
    double* freq;
    double* em_post;
    int32_t* call;
    double* var_phred;
    int32_t* em_iters;
};

__device__ __forceinline__ void genotype_pair(int g, int H, int& h1, int& h2) {
    int i = 0, rem = g;  // order of cgenotype.pyx:193-218: (0,0),(0,1)..(0,H-1),(1,1)..
    while (rem >= H - i) {
        rem -= H - i;
        ++i;
    }
    h1 = i;
    h2 = i + rem;
}

// The per-read mixture has four branches (cgenotype.pyx:164-180); only the last one needs exp/log and
// it is rare, but with one thread per genotype a warp pays for it whenever ANY of its lanes takes it.
// So the (genotype, read) pairs that need it are first listed (count -> prefix -> fill), evaluated
// densely by all threads, and the ordered per-genotype sums then pick the finished terms up in read
// order - every sum adds the same values in the same order as the reference.
constexpr int kMidCap = 512;    // listed terms per round of 64 genotypes; overflow is computed in place

__device__ __forceinline__ double mix_mid(double a, double c) { return log(0.5 * (exp(a) + exp(c))); }
// 0 = homozygous or |a-c| <= 1e-3 (term a), 1 = |a-c| >= 3 (log(1/2) + max), 2 = full mixture
__device__ __forceinline__ int mix_class(bool hom, double a, double c) {
    if (hom) return 0;
    const double d = fabs(a - c);
    if (d >= 3) return 1;
    if (d <= 1e-3) return 0;
    return 2;
}

__global__ void __launch_bounds__(64) k_genotype(DevBatch b, const double* __restrict__ ll, PopOut out, int wi_base) {
    const int wi = wi_base + blockIdx.x;
    const int nInd = b.n_individuals;
    const int w = wi / nInd, i = wi % nInd;
    const int H = b.win_hap_off[w + 1] - b.win_hap_off[w];
    const int G = H * (H + 1) / 2;
    const int Hmax = out.max_haps, Gmax = Hmax * (Hmax + 1) / 2;
    const int T = (int)(b.wi_slot_off[wi + 1] - b.wi_slot_off[wi]);
    const int ngood = b.wi_n_good[wi];
    const double* L = ll + b.ll_off[wi];
    double* gl = out.gl + ((size_t)w * nInd + i) * Gmax;
    __shared__ double s_max[64];
    __shared__ double s_mid[kMidCap];
    __shared__ uint32_t s_ent[kMidCap];    // genotype (local to the round) << 16 | read... as two u16 when T < 65536
    __shared__ int s_off[65];
    const int tid = threadIdx.x;

    double mymax = -1e7;  // cpopulation.pyx:288
    for (int g0 = 0; g0 < G; g0 += 64) {   // rounds of 64 genotypes (uniform trip count)
        const int g = g0 + tid;
        int h1 = 0, h2 = 0;
        const bool act = g < G && ngood != 0;
        if (g < G) genotype_pair(g, H, h1, h2);
        const double* a1 = L + (size_t)h1 * T;
        const double* a2 = L + (size_t)h2 * T;
        const bool hom = (h1 == h2);
        // pass 1: how many full-mixture terms does this genotype have
        int n_mid = 0;
        if (act && !hom)
            for (int t = 0; t < T; ++t) n_mid += (mix_class(false, a1[t], a2[t]) == 2);
        __syncthreads();   // previous round's readers of s_off / s_mid are done
        s_off[tid + 1] = n_mid;
        if (tid == 0) s_off[0] = 0;
        __syncthreads();
        if (tid == 0)
            for (int k = 1; k <= 64; ++k) s_off[k] += s_off[k - 1];
        __syncthreads();
        const int off = s_off[tid], total = min(s_off[64], kMidCap);
        // pass 2: list them (read order inside a genotype)
        if (n_mid > 0) {
            int k = off;
            for (int t = 0; t < T && k < kMidCap; ++t)
                if (mix_class(false, a1[t], a2[t]) == 2) s_ent[k++] = ((uint32_t)tid << 24) | (uint32_t)t;
        }
        __syncthreads();
        // pass 3: all threads evaluate the listed terms
        for (int e = tid; e < total; e += 64) {
            const uint32_t en = s_ent[e];
            int e1, e2;
            genotype_pair(g0 + (int)(en >> 24), H, e1, e2);
            const int t = (int)(en & 0xFFFFFFu);
            s_mid[e] = mix_mid(L[(size_t)e1 * T + t], L[(size_t)e2 * T + t]);
        }
        __syncthreads();
        // pass 4: the ordered sums (cgenotype.pyx:151-180)
        if (g < G) {
            double v = 1.0, gof = 0.0;
            if (ngood != 0) {
                double like = 0.0, gsum = 0.0;
                int k = off;
                for (int t = 0; t < T; ++t) {
                    const double a = a1[t], c = a2[t];
                    const double la = kLog10E * a, lc = kLog10E * c;
                    gsum += la > lc ? la : lc;
                    const int cls = mix_class(hom, a, c);
                    if (cls == 0) {
                        like += a;
                    } else if (cls == 1) {
                        like += (kLogHalf + (a > c ? a : c));
                    } else {
                        like += (k < kMidCap) ? s_mid[k] : mix_mid(a, c);
                        ++k;
                    }
                }
                v = like;
                gof = (-10 * gsum) / ngood;  // cgenotype.pyx:182-183
                if (v > mymax) mymax = v;
            }
            gl[g] = v;
            if (out.gof) out.gof[((size_t)w * Gmax + g) * nInd + i] = gof;
        }
    }
    for (int g = G + tid; g < Gmax; g += 64) {
        gl[g] = 0.0;
        if (out.gof) out.gof[((size_t)w * Gmax + g) * nInd + i] = 0.0;
    }
    if (out.hap_like) {
        for (int h = tid; h < Hmax; h += 64) {
            double sum = 0.0;
            if (h < H)
                for (int t = 0; t < T; ++t) sum += kLog10E * L[(size_t)h * T + t];
            out.hap_like[((size_t)w * nInd + i) * Hmax + h] = sum;
        }
    }
    s_max[tid] = mymax;
    __syncthreads();
    for (int o = 32; o > 0; o >>= 1) {
        if (tid < o) s_max[tid] = s_max[tid] > s_max[tid + o] ? s_max[tid] : s_max[tid + o];
        __syncthreads();
    }
    const double maxll = s_max[0];
    if (tid == 0 && out.gl_log_max) out.gl_log_max[wi] = maxll;
    for (int g = tid; g < G; g += 64) {  // cpopulation.pyx:304-309
        double v = 1.0;
        if (ngood != 0) {
            v = exp(gl[g] - maxll);
            if (!(v > 1e-300)) v = 1e-300;
        }
        gl[g] = v;
    }
}

// ---------------------------------------------------------------------------------------------
// k_population: one block per window, a THREAD PER INDIVIDUAL (many-sample batches; k_population_few
// handles batches with a handful of individuals).  EM (cpopulation.pyx:384-457, 678-703), genotype calls
// (:623-676), variant posteriors (:459-594).  blockDim.x = NT threads (a multiple of 32, up to 512);
// thread t owns individuals t, t+NT, ...
//
// Everything per individual is independent and runs across the threads in the reference's own order of
// operations.  The sums ACROSS individuals are floating-point chains whose order the reference fixes -
//   newFreqs[k]:  individuals ascending, inside one its genotypes ascending, `+= csr` once per haplotype slot
//                 (twice for the homozygous genotype)                                  (cpopulation.pyx:436-447)
//   sum of log P(variant) / log P(no variant) over the individuals, ascending          (:546-581)
// - and a different association changes the last bits, which can flip the EM's stop test (one iteration more or
// fewer) or a rounded phred value.  So each chain is walked by ONE thread in exactly that order (haplotype k's chain by
// thread k; the posterior's by thread 0 over values the other threads have laid out in shared memory).  With 2000
// individuals and 8 haplotypes that is 18,000 dependent additions per EM iteration per window - microseconds next to
// the alignment work of the same window - and it makes call / em_iters / var_phred equal to the reference's bit for bit.
// Shared memory: 3 * Hmax doubles.
// ---------------------------------------------------------------------------------------------
constexpr int kPopMaxThreads = 512;

__global__ void __launch_bounds__(kPopMaxThreads) k_population(DevBatch b, PopOut out, double* __restrict__ em_scratch,
                                                               int max_iters, int use_em, int nthr_em, int w_base) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int w = w_base + blockIdx.x;
    const int nInd = b.n_individuals;
    const int NT = blockDim.x;
    const int H = b.win_hap_off[w + 1] - b.win_hap_off[w];
    const int G = H * (H + 1) / 2;
    const int Hmax = out.max_haps, Gmax = Hmax * (Hmax + 1) / 2;
    double* s_freq = (double*)smem;             // [Hmax]
    double* s_new = s_freq + Hmax;              // [Hmax]
    double* s_fp = s_new + Hmax;                // [Hmax] frequencies without one variant's haplotypes
    __shared__ double s_lv[kPopMaxThreads], s_ln[kPopMaxThreads];   // one tile of per-individual log terms
    __shared__ double s_change, s_slv;
    __shared__ int s_nwith;
    const int tid = threadIdx.x;
    (void)nthr_em;
    const double* gl = out.gl + (size_t)w * nInd * Gmax;
    double* emp = (out.em_post ? out.em_post : em_scratch) + (size_t)w * nInd * Gmax;
    const int32_t* ngood = b.wi_n_good + (size_t)w * nInd;

    const double eps = fmin(1e-3, 1.0 / (nInd * 2 * 2));  // cpopulation.pyx:684
    for (int k = tid; k < Hmax; k += NT) s_freq[k] = k < H ? 1.0 / H : 0.0;
    if (tid == 0) {
        s_nwith = 0;
        s_change = eps + 1;
    }
    __syncthreads();
    int my_with = 0;
    for (int i = tid; i < nInd; i += NT) {
        if (ngood[i] == 0) {
            for (int g = 0; g < Gmax; ++g) emp[(size_t)i * Gmax + g] = 0.0;
        } else {
            ++my_with;
            for (int g = G; g < Gmax; ++g) emp[(size_t)i * Gmax + g] = 0.0;
        }
    }
    if (my_with) atomicAdd(&s_nwith, my_with);   // an integer count: order does not matter
    __syncthreads();
    const int n_with = s_nwith;                  // individuals with reads
    int iters = 0;
    while (s_change > eps && iters < max_iters) {  // uniform: s_change only changes between barriers
        // E step, one individual per thread (cpopulation.pyx:407-430)
        for (int i = tid; i < nInd; i += NT) {
            if (ngood[i] == 0) continue;
            const double* gli = gl + (size_t)i * Gmax;
            double* csr = emp + (size_t)i * Gmax;
            double sum = 0.0;
            int g = 0;
            for (int s = 0; s < H; ++s)
                for (int r = s; r < H; ++r, ++g) {
                    const double v = gli[g] * s_freq[s] * s_freq[r] * (1 + (r != s));
                    csr[g] = v;
                    sum += v;
                }
            if (sum > 0.0)
                for (g = 0; g < G; ++g) csr[g] /= sum;
        }
        __syncthreads();   // the rows written above are read by other threads of this block below
        // M step: haplotype k's chain by thread k, in the reference's order (see the header comment)
        for (int k = tid; k < H; k += NT) {
            double acc = 0.0;
            for (int i = 0; i < nInd; ++i) {
                if (ngood[i] == 0) continue;
                const double* csr = emp + (size_t)i * Gmax;
                int g = k;                                   // g(0, k)
                for (int s = 0; s < k; ++s) {                // genotypes (s, k), s < k: g(s+1, k) = g(s, k) + H - s - 1
                    acc += csr[g];
                    g += H - s - 1;
                }
                const double hom = csr[g];                   // g(k, k): both slots are haplotype k
                acc += hom;
                acc += hom;
                for (int r = k + 1; r < H; ++r) acc += csr[g + (r - k)];
            }
            s_new[k] = acc;
        }
        __syncthreads();
        if (tid == 0) {   // cpopulation.pyx:449-455
            double m = 0.0;
            for (int k = 0; k < H; ++k) {
                const double nf = s_new[k] / (2 * n_with);
                const double ch = fabs(s_freq[k] - nf);
                if (ch > m) m = ch;
                s_freq[k] = nf;
            }
            s_change = m;
        }
        ++iters;
        __syncthreads();
    }
    if (out.freq)
        for (int k = tid; k < Hmax; k += NT) out.freq[(size_t)w * Hmax + k] = s_freq[k];
    if (tid == 0 && out.em_iters) out.em_iters[w] = iters;
    // callGenotypes, cpopulation.pyx:623-676: first strict maximum
    if (out.call) {
        for (int i = tid; i < nInd; i += NT) {
            int bestg = -1;
            double bestv = 0.0;
            if (ngood[i] != 0) {
                const double* src = (use_em == 1 ? emp : gl) + (size_t)i * Gmax;
                for (int g = 0; g < G; ++g)
                    if (bestg == -1 || src[g] > bestv) {
                        bestv = src[g];
                        bestg = g;
                    }
            }
            out.call[(size_t)w * nInd + i] = bestg;
        }
    }
    // calculatePosterior, cpopulation.pyx:459-594.  Per individual P(variant) does not depend on the variant, so the
    // sum of its logs is formed once (pass v = -1); per variant the individuals' log P(no variant) are laid out a tile
    // at a time and thread 0 adds them in order.
    if (out.var_phred && b.max_variants > 0 && b.win_n_var) {
        const int nvar = b.win_n_var[w];
        const uint64_t* masks = b.hap_var_mask + b.win_hap_off[w];
        for (int v = -1; v < b.max_variants; ++v) {
            double ph = 0.0;
            if (v < nvar && (v >= 0 || nvar > 0)) {   // uniform over the block
                __syncthreads();
                if (tid == 0 && v >= 0) {
                    double sumf = 0.0;
                    for (int k = 0; k < H; ++k) {
                        if (!((masks[k] >> v) & 1ull)) {
                            s_fp[k] = s_freq[k];
                            sumf += s_freq[k];
                        } else {
                            s_fp[k] = 0.0;
                        }
                    }
                    if (sumf > 0)
                        for (int k = 0; k < H; ++k) s_fp[k] /= sumf;
                }
                __syncthreads();
                const double* fq = v < 0 ? s_freq : s_fp;
                double chain = 0.0;   // thread 0 only
                for (int base = 0; base < nInd; base += NT) {
                    const int i = base + tid;
                    double term = 0.0;
                    bool have = false;
                    if (i < nInd && ngood[i] != 0) {
                        const double* gli = gl + (size_t)i * Gmax;
                        double p = 0.0;
                        int g = 0;
                        for (int r = 0; r < H; ++r)
                            for (int s = r; s < H; ++s, ++g) {
                                const double factor = (r != s) ? 2.0 : 1.0;
                                p += (factor * fq[r] * fq[s] * gli[g]);
                            }
                        term = p > 0 ? log(p) : -708.0;
                        have = true;
                    }
                    s_lv[tid] = term;
                    s_ln[tid] = have ? 1.0 : 0.0;
                    __syncthreads();
                    if (tid == 0) {
                        const int lim = min(NT, nInd - base);
                        for (int t = 0; t < lim; ++t)
                            if (s_ln[t] != 0.0) chain += s_lv[t];
                    }
                    __syncthreads();
                }
                if (tid == 0) {
                    if (v < 0) {
                        s_slv = chain;
                    } else {
                        double ratio = exp(chain - s_slv);
                        if (!(ratio > 1e-300)) ratio = 1e-300;
                        const double prior = b.var_prior[(size_t)w * b.max_variants + v];
                        ph = round(-10.0 * (log10(ratio * (1.0 - prior)) - log10(prior + ratio * (1.0 - prior))));
                    }
                }
            }
            if (tid == 0 && v >= 0) out.var_phred[(size_t)w * b.max_variants + v] = ph;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// k_population_few: same model as k_population for batches with FEW individuals, where a thread per
// individual leaves the GPU idle.  One WARP per window; inside an individual the lanes work across
// genotypes / haplotypes, and every floating-point sum still runs in the reference's order:
//   * csr[g] = GL * f_s * f_r * (1 + [r != s])        lanes over g (independent products)
//   * sum over g                                        lane 0, ascending g        (cpopulation.pyx:420-427)
//   * csr[g] /= sum                                     lanes over g
//   * newFreq[k] += csr[g] for g containing k           lane k, ascending g, twice when homozygous
//                                                       (cpopulation.pyx:431-440: part[s] += v; part[r] += v)
// Individuals are visited in order, so newFreq accumulates exactly as in the reference.
// Shared memory per warp: Gmax doubles (csr) + 2*Hmax doubles.
// ---------------------------------------------------------------------------------------------
constexpr int kPopFewWarps = 4;

__global__ void __launch_bounds__(32 * kPopFewWarps) k_population_few(DevBatch b, PopOut out, double* __restrict__ em_scratch,
                                                                  int max_iters, int use_em, int w_base, int w_end) {
    extern __shared__ __align__(16) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int w = w_base + blockIdx.x * kPopFewWarps + warp;
    if (w >= w_end) return;
    const unsigned FULL = 0xFFFFFFFFu;
    const int nInd = b.n_individuals;
    const int H = b.win_hap_off[w + 1] - b.win_hap_off[w];
    const int G = H * (H + 1) / 2;
    const int Hmax = out.max_haps, Gmax = Hmax * (Hmax + 1) / 2;
    double* s_csr = (double*)smem + (size_t)warp * (Gmax + 2 * Hmax);
    double* s_freq = s_csr + Gmax;
    double* s_new = s_freq + Hmax;
    const double* gl = out.gl + (size_t)w * nInd * Gmax;
    double* emp = (out.em_post ? out.em_post : em_scratch) + (size_t)w * nInd * Gmax;
    const int32_t* ngood = b.wi_n_good + (size_t)w * nInd;

    const double eps = fmin(1e-3, 1.0 / (nInd * 2 * 2));  // cpopulation.pyx:684
    for (int k = lane; k < Hmax; k += 32) s_freq[k] = k < H ? 1.0 / H : 0.0;
    for (int i = 0; i < nInd; ++i) {
        if (ngood[i] == 0)
            for (int g = lane; g < Gmax; g += 32) emp[(size_t)i * Gmax + g] = 0.0;
        for (int g = G + lane; g < Gmax; g += 32) emp[(size_t)i * Gmax + g] = 0.0;
    }
    __syncwarp();
    double change = eps + 1;
    int iters = 0;
    while (change > eps && iters < max_iters) {
        for (int k = lane; k < H; k += 32) s_new[k] = 0.0;
        int n_with = 0;
        for (int i = 0; i < nInd; ++i) {
            if (ngood[i] == 0) continue;
            ++n_with;
            const double* gli = gl + (size_t)i * Gmax;
            double* csr = emp + (size_t)i * Gmax;
            __syncwarp();
            for (int g = lane; g < G; g += 32) {
                int sI, rI;
                genotype_pair(g, H, sI, rI);
                s_csr[g] = gli[g] * s_freq[sI] * s_freq[rI] * (1 + (rI != sI));
            }
            __syncwarp();
            double sum = 0.0;
            if (lane == 0)
                for (int g = 0; g < G; ++g) sum += s_csr[g];
            sum = __shfl_sync(FULL, sum, 0);
            for (int g = lane; g < G; g += 32) {
                double v = s_csr[g];
                if (sum > 0.0) v /= sum;
                s_csr[g] = v;
                csr[g] = v;
            }
            __syncwarp();
            for (int k = lane; k < H; k += 32) {   // haplotype k: its genotypes in ascending g
                double acc = s_new[k];
                int g = 0;
                for (int sI = 0; sI < H; ++sI)
                    for (int rI = sI; rI < H; ++rI, ++g) {
                        if (sI == k) acc += s_csr[g];
                        if (rI == k) acc += s_csr[g];
                    }
                s_new[k] = acc;
            }
        }
        __syncwarp();
        double mych = 0.0;
        for (int k = lane; k < H; k += 32) {
            double nf = s_new[k];
            if (n_with > 0) nf = nf / (2 * n_with); else nf = s_freq[k];
            const double ch = fabs(s_freq[k] - nf);
            if (ch > mych) mych = ch;
            s_new[k] = nf;
        }
        for (int o = 16; o > 0; o >>= 1) {
            const double other = __shfl_xor_sync(FULL, mych, o);
            mych = other > mych ? other : mych;
        }
        __syncwarp();
        for (int k = lane; k < H; k += 32) s_freq[k] = s_new[k];
        __syncwarp();
        change = mych;
        ++iters;
    }
    if (out.freq)
        for (int k = lane; k < Hmax; k += 32) out.freq[(size_t)w * Hmax + k] = s_freq[k];
    if (lane == 0 && out.em_iters) out.em_iters[w] = iters;
    // callGenotypes, cpopulation.pyx:623-676: first strict maximum
    if (out.call) {
        for (int i = lane; i < nInd; i += 32) {
            int bestg = -1;
            double bestv = 0.0;
            if (ngood[i] != 0) {
                const double* src = (use_em == 1 ? emp : gl) + (size_t)i * Gmax;
                for (int g = 0; g < G; ++g)
                    if (bestg == -1 || src[g] > bestv) {
                        bestv = src[g];
                        bestg = g;
                    }
            }
            out.call[(size_t)w * nInd + i] = bestg;
        }
    }
    // calculatePosterior, cpopulation.pyx:459-594: one lane per variant, genotypes in order
    if (out.var_phred && b.max_variants > 0 && b.win_n_var) {
        const int nvar = b.win_n_var[w];
        const uint64_t* masks = b.hap_var_mask + b.win_hap_off[w];
        __syncwarp();
        for (int v = lane; v < b.max_variants; v += 32) {
            double ph = 0.0;
            if (v < nvar) {
                double sumf = 0.0;
                for (int k = 0; k < H; ++k)
                    if (!((masks[k] >> v) & 1ull)) sumf += s_freq[k];
                auto fp = [&](int k) -> double {   // frequencies with the variant's haplotypes removed, renormalised
                    if ((masks[k] >> v) & 1ull) return 0.0;
                    return sumf > 0 ? s_freq[k] / sumf : s_freq[k];
                };
                double slv = 0.0, sln = 0.0;
                for (int i = 0; i < nInd; ++i) {
                    if (ngood[i] == 0) continue;
                    const double* gli = gl + (size_t)i * Gmax;
                    double pv = 0.0, pn = 0.0;
                    int g = 0;
                    for (int r = 0; r < H; ++r) {
                        const double fr = fp(r);
                        for (int sI = r; sI < H; ++sI, ++g) {
                            const double l = gli[g];
                            const double factor = (r != sI) ? 2.0 : 1.0;
                            pv += (factor * s_freq[r] * s_freq[sI] * l);
                            pn += (factor * fr * fp(sI) * l);
                        }
                    }
                    slv += pv > 0 ? log(pv) : -708.0;
                    sln += pn > 0 ? log(pn) : -708.0;
                }
                double ratio = exp(sln - slv);
                if (!(ratio > 1e-300)) ratio = 1e-300;
                const double prior = b.var_prior[(size_t)w * b.max_variants + v];
                ph = round(-10.0 * (log10(ratio * (1.0 - prior)) - log10(prior + ratio * (1.0 - prior))));
            }
            out.var_phred[(size_t)w * b.max_variants + v] = ph;
        }
    }
}


// ---------------------------------------------------------------------------------------------
// k_site_genotypes (scope row N4): computeGenotypeCallAndLikelihoods (src/cython/vcfutils.pyx:163-334)
// and the per-sample derivations of outputCallToVCF (vcfutils.pyx:491-548), one thread per
// (site, individual); allele pairs and genotypes are visited in the reference's order, so every sum
// adds the same values in the same order.
// ---------------------------------------------------------------------------------------------
struct SiteIn {
    int32_t n_sites, n_individuals, max_haps, min_posterior;
    const int32_t* site_win;
    const int32_t* site_var_off;
    const int32_t* site_var;
    const int64_t* site_hap_off;
    const uint8_t* hap_is_ref;
    const int32_t* win_hap_off;
    const uint64_t* hap_var_mask;
    const int32_t* wi_n_good;
    const double* gl;
    const double* gof;
    const double* freq;
};
struct SiteOutDev {
    int32_t max_pairs;
    int32_t* phased;
    double* lik;
    double* post;
    int32_t* phred;
    double* gof;
    int32_t* gt;
    double* gl_log10;
};

__device__ __forceinline__ double py_max(double a, double b) { return b > a ? b : a; }   // Python max(a, b)
__device__ __forceinline__ double py_min(double a, double b) { return b < a ? b : a; }

__global__ void __launch_bounds__(128) k_site_genotypes(SiteIn in, SiteOutDev out) {
    const int64_t o = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int nInd = in.n_individuals;
    if (o >= (int64_t)in.n_sites * nInd) return;
    const int s = (int)(o / nInd), i = (int)(o % nInd);
    const int w = in.site_win[s];
    const int h0 = in.win_hap_off[w], H = in.win_hap_off[w + 1] - h0;
    const int Hm = in.max_haps, Gm = Hm * (Hm + 1) / 2;
    const int P = out.max_pairs;
    const int nV = in.site_var_off[s + 1] - in.site_var_off[s];
    const int32_t* vars = in.site_var + in.site_var_off[s];
    const uint8_t* is_ref = in.hap_is_ref + in.site_hap_off[s];
    const uint64_t* masks = in.hap_var_mask + h0;
    const double* freq = in.freq + (size_t)w * Hm;
    const double* gl = in.gl + ((size_t)w * nInd + i) * Gm;
    if (out.lik)
        for (int k = 0; k < P; ++k) out.lik[o * P + k] = 0.0;
    if (in.wi_n_good[(size_t)w * nInd + i] == 0) {   // vcfutils.pyx:497-499
        if (out.phased) out.phased[o * 2] = out.phased[o * 2 + 1] = -1;
        if (out.post) out.post[o * 3] = out.post[o * 3 + 1] = out.post[o * 3 + 2] = 0.0;
        if (out.phred) out.phred[o * 3] = out.phred[o * 3 + 1] = out.phred[o * 3 + 2] = 0;
        if (out.gof) out.gof[o] = 0.0;
        if (out.gt) out.gt[o * 2] = out.gt[o * 2 + 1] = -1;
        if (out.gl_log10) out.gl_log10[o * 3] = out.gl_log10[o * 3 + 1] = out.gl_log10[o * 3 + 2] = 0.0;
        return;
    }
    double sum_lik = 0.0, best_gof = 1e6, best_lik = -1.0, nonref = 0.0, ref = 0.0, phased_max = -1e6;
    double liks[3] = {0.0, 0.0, 0.0}, max_lik = 0.0;
    int ph1 = -1, ph2 = -1, pair = 0;
    bool have_max = false;
    for (int i1 = 0; i1 <= nV; ++i1)
        for (int i2 = 0; i2 <= i1; ++i2, ++pair) {
            double marg = 0.0;
            int g = 0;
            for (int a = 0; a < H; ++a)
                for (int c = a; c < H; ++c, ++g) {
                    const bool ref1 = is_ref[a] != 0, ref2 = is_ref[c] != 0;
                    const double factor = (a != c) ? 2.0 : 1.0;
                    bool v1h1 = false, v1h2 = false, v2h1 = false, v2h2 = false, match;
                    if (i1 == 0 && i2 == 0) {
                        match = ref1 && ref2;
                    } else if (i2 == 0) {
                        v1h1 = (masks[a] >> vars[i1 - 1]) & 1ull;
                        v1h2 = (masks[c] >> vars[i1 - 1]) & 1ull;
                        match = (ref2 && v1h1) || (ref1 && v1h2);
                    } else {
                        v1h1 = (masks[a] >> vars[i1 - 1]) & 1ull;
                        v1h2 = (masks[c] >> vars[i1 - 1]) & 1ull;
                        v2h1 = (masks[a] >> vars[i2 - 1]) & 1ull;
                        v2h2 = (masks[c] >> vars[i2 - 1]) & 1ull;
                        match = (v1h1 && v2h2) || (v2h1 && v1h2);
                    }
                    if (!match) continue;
                    const double cur = nInd > 25 ? (factor * freq[a] * freq[c] * gl[g]) : (factor * gl[g]);
                    marg += cur;
                    if (cur > phased_max) {   // phase by the maximum-likelihood genotype, :276-316
                        phased_max = cur;
                        if (i1 == 0 && i2 == 0) {
                            ph1 = i1;
                            ph2 = i2;
                        } else if (i2 == 0 && i1 != 0) {
                            if (v1h1) {
                                ph1 = i1;
                                ph2 = i2;
                            } else if (v1h2) {
                                ph1 = i2;
                                ph2 = i1;
                            }
                        } else if (i2 == i1 && i1 > 0) {
                            ph1 = i1;
                            ph2 = i2;
                        } else if (i2 > 0 && i1 > 0 && i2 != i1) {
                            if (v1h1 && v2h2) {
                                ph1 = i1;
                                ph2 = i2;
                            } else if (v1h2 && v2h1) {
                                ph1 = i2;
                                ph2 = i1;
                            }
                        }
                    }
                    const double gf = in.gof[((size_t)w * Gm + g) * nInd + i];
                    if (gf < best_gof) best_gof = gf;
                }
            if (marg > best_lik) best_lik = marg;
            if ((i1 == 1 && i2 == 0) || (i1 == 1 && i2 == 1)) nonref += marg;
            else if (i1 == 0 && i2 == 0) ref += marg;
            sum_lik += marg;
            if (out.lik && pair < P) out.lik[o * P + pair] = marg;
            if (pair < 3) liks[pair] = marg;
            if (!have_max || marg > max_lik) {
                max_lik = marg;
                have_max = true;
            }
        }
    const double gpost = best_lik / sum_lik, npost = nonref / sum_lik, rpost = ref / sum_lik;
    const int q_g = (int)py_min(99, round(-10.0 * log10(py_max(1e-10, 1.0 - gpost))));
    const int q_n = (int)py_min(99, round(-10.0 * log10(py_max(1e-10, 1.0 - npost))));
    const int q_r = (int)py_min(99, round(-10.0 * log10(py_max(1e-10, 1.0 - rpost))));
    int gt1 = ph1, gt2 = ph2;
    double gl3[3] = {-1.0, -1.0, -1.0};
    if (nV == 1) {   // vcfutils.pyx:518-532
        if (q_n < in.min_posterior && q_r < in.min_posterior) gt1 = gt2 = -1;
        else if (q_n < in.min_posterior) gt1 = gt2 = 0;
        for (int k = 0; k < 3; ++k) gl3[k] = log10(py_max(liks[k] / max_lik, 1e-300));
    }
    if (out.phased) {
        out.phased[o * 2] = ph1;
        out.phased[o * 2 + 1] = ph2;
    }
    if (out.post) {
        out.post[o * 3] = gpost;
        out.post[o * 3 + 1] = npost;
        out.post[o * 3 + 2] = rpost;
    }
    if (out.phred) {
        out.phred[o * 3] = q_g;
        out.phred[o * 3 + 1] = q_n;
        out.phred[o * 3 + 2] = q_r;
    }
    if (out.gof) out.gof[o] = best_gof;
    if (out.gt) {
        out.gt[o * 2] = gt1;
        out.gt[o * 2 + 1] = gt2;
    }
    if (out.gl_log10)
        for (int k = 0; k < 3; ++k) out.gl_log10[o * 3 + k] = gl3[k];
}

}  // namespace plb
