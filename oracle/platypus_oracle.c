/*
 * platypus_oracle.c — CPU restatement of the Platypus read-vs-haplotype likelihood path.
 *
 * TEST INFRASTRUCTURE ONLY (see platypus_oracle.h).  Written from the behaviour of the
 * reference, function by function; every function cites the reference lines it follows
 * (paths relative to the reference checkout).  Plain C, double arithmetic in the same
 * order as the reference so that results can be compared tightly.
 */
#include "platypus_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define KMER 7
#define HASH_SIZE 16384          /* 4^7, calign.pyx:25-27 */
#define BAND 16
#define BIG (1 << 28)

static const double M_LTOT = -0.23025850929940459;  /* chaplotype.pyx / calign.pyx:31 */
static const double LOG10E = 0.43429448190325182;    /* cgenotype.pyx:24 */
static const double LOG_HALF = -0.69314718055994529; /* cgenotype.pyx:28 */

static plo_align_fn g_align_fn = 0;
static int g_align_traceback = 0;

void plo_set_align_fn(plo_align_fn fn, int traceback) {
    g_align_fn = fn;
    g_align_traceback = traceback;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * L1 — banded affine-gap min-plus alignment, src/c/align.c:77-586.
 *
 * The reference walks anti-diagonals with 8 SSE lanes; restated here cell by cell.
 * x indexes the haplotype segment (0 .. read_len+14), y the read, diagonal d = x - y in
 * [0,15].  Per cell:
 *   sub  = 0 if hap[x]=='N' or hap[x]==read[y], else qual[y]          (align.c:175-178,314-318)
 *   M    = min(M,I,D)(x-1,y-1) + sub, with the y=-1 row treated as 0  (initmask, :244-250)
 *   I    = min(I(x,y-1)+ext, M(x,y-1)+open[x]) + nucprior             (:331-335, 478-484)
 *          at y=0 only even x see a zero M above them (the even-lane init of :244-250)
 *   D    = min(D(x-1,y)+ext, min(M,I)(x-1,y)+open[x])                 (:320-329, 472-476)
 * Result: min over the 16 band cells of the last read row (:261-288, 416-443, 520).
 * -----------------------------------------------------------------------------------------*/
int plo_band_align(const uint8_t* hap, const uint8_t* read, const uint8_t* qual, int read_len,
                   int ext, int nuc, const uint8_t* open) {
    int Mp[BAND], Ip[BAND], Dp[BAND], Mc[BAND], Ic[BAND], Dc[BAND];
    for (int y = 0; y < read_len; ++y) {
        for (int d = 0; d < BAND; ++d) {
            int x = y + d;
            int sub = (hap[x] == 'N' || hap[x] == read[y]) ? 0 : (int)qual[y];
            int diag = (y == 0) ? 0 : imin(Mp[d], imin(Ip[d], Dp[d]));
            int m = diag + sub;
            int ins;
            if (y == 0) {
                ins = (x % 2 == 0) ? (int)open[x] + nuc : BIG;
            } else if (d + 1 < BAND) {
                ins = imin(Ip[d + 1] + ext, Mp[d + 1] + (int)open[x]) + nuc;
            } else {
                ins = BIG;
            }
            int del;
            if (d >= 1) {
                del = imin(Dc[d - 1] + ext, imin(Mc[d - 1], Ic[d - 1]) + (int)open[x]);
            } else {
                del = BIG;
            }
            Mc[d] = imin(m, BIG);
            Ic[d] = imin(ins, BIG);
            Dc[d] = imin(del, BIG);
        }
        memcpy(Mp, Mc, sizeof Mp);
        memcpy(Ip, Ic, sizeof Ip);
        memcpy(Dp, Dc, sizeof Dp);
    }
    int best = BIG;
    for (int d = 0; d < BAND; ++d) best = imin(best, imin(Mp[d], imin(Ip[d], Dp[d])));
    return best;
}

/* ------------------------------------------------------------------------------------------
 * L1 with traceback, src/c/align.c:344-365, 493-515 (back-pointers) and :523-577 (walk).
 *
 * With traceback on, every int16 value of the reference carries the label of its own state in
 * its two low bits (M=0, I=1, D=3; scores are x4 so sums keep the label of the operand) and
 * every `min` compares value AND label: equal scores resolve to the smaller label (M < I < D).
 * The label that survives a cell's min is that state's back-pointer.  Restated here with
 * key = 4*score + label.  The end cell is the first diagonal (ascending) holding the
 * smallest key of the last read row (:261-288, 416-443 compare labelled values).
 * Out-of-band / ramp-up values are "infinite" and never lie on the chosen path, so their
 * garbage pointers are not reproduced.
 * -----------------------------------------------------------------------------------------*/
#define LBL_M 0
#define LBL_I 1
#define LBL_D 3
#define KBIG ((int64_t)1 << 40)

typedef struct { int64_t m, i, d; } Keys3;      /* 4*score + own label */

static inline int64_t kmin(int64_t a, int64_t b) { return a < b ? a : b; }
static inline int64_t relabel(int64_t k, int lbl) { return k >= KBIG ? KBIG + lbl : ((k >> 2) << 2) + lbl; }

/* Computes the banded matrix keeping the three back-pointers of every cell in ptr[y*16+d]
 * (bits 0-1 M, 2-3 I, 6-7 D, as align.c:346-348 packs them).  Returns the end diagonal, the
 * state to start the walk in and the score. */
static int band_forward_tb(const uint8_t* hap, const uint8_t* read, const uint8_t* qual, int L, int ext, int nuc,
                           const uint8_t* open, uint8_t* ptr, int* end_diag, int* end_state) {
    Keys3 prev[BAND], cur[BAND];
    for (int y = 0; y < L; ++y) {
        for (int d = 0; d < BAND; ++d) {
            int x = y + d;
            int sub = (hap[x] == 'N' || hap[x] == read[y]) ? 0 : (int)qual[y];
            int64_t diag = (y == 0) ? (int64_t)LBL_M : kmin(prev[d].m, kmin(prev[d].i, prev[d].d));
            int pm = (int)(diag & 3);
            int64_t m = diag >= KBIG ? KBIG : diag + 4 * sub;
            int64_t ins;
            if (y == 0) ins = (x % 2 == 0) ? (int64_t)LBL_M + 4 * ((int)open[x] + nuc) : KBIG;
            else if (d + 1 < BAND) {
                int64_t a = prev[d + 1].i >= KBIG ? KBIG : prev[d + 1].i + 4 * ext;
                int64_t bq = prev[d + 1].m >= KBIG ? KBIG : prev[d + 1].m + 4 * (int)open[x];
                ins = kmin(a, bq);
                if (ins < KBIG) ins += 4 * nuc;
            } else ins = KBIG;
            int pi = (int)(ins & 3);
            int64_t del;
            if (d >= 1) {
                int64_t a = cur[d - 1].d >= KBIG ? KBIG : cur[d - 1].d + 4 * ext;
                int64_t mi = kmin(cur[d - 1].m, cur[d - 1].i);
                int64_t bq = mi >= KBIG ? KBIG : mi + 4 * (int)open[x];
                del = kmin(a, bq);
            } else del = KBIG;
            int pd = (int)(del & 3);
            ptr[y * BAND + d] = (uint8_t)(pm | (pi << 2) | (pd << 6));
            cur[d].m = relabel(m, LBL_M);
            cur[d].i = relabel(ins, LBL_I);
            cur[d].d = relabel(del, LBL_D);
        }
        memcpy(prev, cur, sizeof prev);
    }
    int64_t best = KBIG + 8;
    int bd = 0;
    for (int d = 0; d < BAND; ++d) {
        int64_t k = kmin(prev[d].m, kmin(prev[d].i, prev[d].d));
        if (k < best) { best = k; bd = d; }
    }
    *end_diag = bd;
    *end_state = (int)(best & 3);
    return (int)(best >> 2);
}

int plo_band_align_tb(const uint8_t* hap, const uint8_t* read, const uint8_t* qual, int L, int ext, int nuc,
                      const uint8_t* open, char* aln1, char* aln2, int* firstpos) {
    uint8_t* ptr = (uint8_t*)malloc((size_t)L * BAND + 1);
    int ed = 0, state = 0;
    int score = band_forward_tb(hap, read, qual, L, ext, nuc, open, ptr, &ed, &state);
    /* walk, align.c:523-577: (cx, cy) is the cell whose symbol pair is emitted next */
    int cx = L - 1 + ed, cy = L - 1, n = 0;
    while (cy >= 0) {
        int d = cx - cy;
        uint8_t p = (d >= 0 && d < BAND) ? ptr[cy * BAND + d] : 0;
        int newstate = (p >> (2 * state)) & 3;
        if (state == LBL_M) { aln1[n] = (char)hap[cx]; aln2[n] = (char)read[cy]; --cx; --cy; }
        else if (state == LBL_I) { aln1[n] = '-'; aln2[n] = (char)read[cy]; --cy; }
        else { aln1[n] = (char)hap[cx]; aln2[n] = '-'; --cx; }
        state = newstate;
        ++n;
    }
    aln1[n] = 0;
    aln2[n] = 0;
    if (firstpos) *firstpos = cx + 1;
    for (int i = 0, j = n - 1; i < j; ++i, --j) { /* :561-571 */
        char t = aln1[i]; aln1[i] = aln1[j]; aln1[j] = t;
        t = aln2[i]; aln2[i] = aln2[j]; aln2[j] = t;
    }
    free(ptr);
    return score;
}

/* src/c/align.c:593-644 calculateFlankScore: cost of the alignment columns whose haplotype
 * coordinate lies in a flank (x < hapFlank or x >= hapLen - hapFlank). */
int plo_flank_score(int hap_len, int hap_flank, const uint8_t* qual, const uint8_t* gap_open, int ext, int nuc,
                    int firstpos, const char* aln1, const char* aln2) {
    char prev = 'M';
    int x = firstpos, y = 0, score = 0;
    for (int i = 0; aln1[i]; ++i) {
        char st = 'M';
        if (aln1[i] == '-') st = 'I';
        if (aln2[i] == '-') st = 'D';
        int in_flank = (x < hap_flank || x >= hap_len - hap_flank);
        if (st == 'M') {
            if (aln1[i] != aln2[i] && in_flank) score += (aln1[i] == 'N') ? 0 : (int)qual[y]; /* n_score/4 = 0 */
            ++x; ++y;
        } else if (st == 'I') {
            if (in_flank) score += (prev == 'I') ? ext + nuc : (int)gap_open[x - 1] + nuc;
            ++y;
        } else {
            if (in_flank) score += (prev == 'D') ? ext : (int)gap_open[x];
            ++x;
        }
        prev = st;
    }
    return score;
}

/* The same two steps fused into one forward pass (what the CUDA path computes): every state
 * carries, next to its key, the in-flank cost of the path its back-pointers describe.  `start`
 * is the offset of the segment in the haplotype, so column x of the band is haplotype
 * position start + x.  An insertion at cell (x, y) is charged at haplotype position x + 1
 * (align.c:621-631: x has already moved past the preceding match).
 * Returns the score; *flank receives calculateFlankScore of the traceback alignment. */
int plo_band_align_flank(const uint8_t* hap, const uint8_t* read, const uint8_t* qual, int L, int ext, int nuc,
                         const uint8_t* open, int start, int hap_len, int hap_flank, int* flank) {
    Keys3 prev[BAND], cur[BAND];
    int fpm[BAND], fpi[BAND], fpd[BAND], fcm[BAND], fci[BAND], fcd[BAND];
    for (int y = 0; y < L; ++y) {
        for (int d = 0; d < BAND; ++d) {
            int x = y + d;
            int hx = start + x;
            int in_x = (hx < hap_flank || hx >= hap_len - hap_flank);
            int in_x1 = (hx + 1 < hap_flank || hx + 1 >= hap_len - hap_flank);
            int sub = (hap[x] == 'N' || hap[x] == read[y]) ? 0 : (int)qual[y];
            /* M */
            int64_t diag;
            int fdiag;
            if (y == 0) { diag = LBL_M; fdiag = 0; }
            else {
                diag = prev[d].m; fdiag = fpm[d];
                if (prev[d].i < diag) { diag = prev[d].i; fdiag = fpi[d]; }
                if (prev[d].d < diag) { diag = prev[d].d; fdiag = fpd[d]; }
            }
            int64_t m = diag >= KBIG ? KBIG : diag + 4 * sub;
            fcm[d] = fdiag + (in_x ? sub : 0);
            /* I */
            int64_t ins;
            int fins = 0;
            if (y == 0) {
                if (x % 2 == 0) { ins = (int64_t)LBL_M + 4 * ((int)open[x] + nuc); fins = in_x1 ? (int)open[x] + nuc : 0; }
                else ins = KBIG;
            } else if (d + 1 < BAND) {
                int64_t a = prev[d + 1].i >= KBIG ? KBIG : prev[d + 1].i + 4 * ext;
                int64_t bq = prev[d + 1].m >= KBIG ? KBIG : prev[d + 1].m + 4 * (int)open[x];
                if (bq <= a) { ins = bq; fins = fpm[d + 1] + (in_x1 ? (int)open[x] + nuc : 0); }   /* label M(0) < I(1) */
                else { ins = a; fins = fpi[d + 1] + (in_x1 ? ext + nuc : 0); }
                if (ins < KBIG) ins += 4 * nuc;
            } else ins = KBIG;
            fci[d] = fins;
            /* D */
            int64_t del;
            int fdel = 0;
            if (d >= 1) {
                int64_t a = cur[d - 1].d >= KBIG ? KBIG : cur[d - 1].d + 4 * ext;
                int64_t mi = cur[d - 1].m;
                int fmi = fcm[d - 1];
                if (cur[d - 1].i < mi) { mi = cur[d - 1].i; fmi = fci[d - 1]; }
                int64_t bq = mi >= KBIG ? KBIG : mi + 4 * (int)open[x];
                if (bq <= a) { del = bq; fdel = fmi + (in_x ? (int)open[x] : 0); }
                else { del = a; fdel = fcd[d - 1] + (in_x ? ext : 0); }
            } else del = KBIG;
            fcd[d] = fdel;
            cur[d].m = relabel(m, LBL_M);
            cur[d].i = relabel(ins, LBL_I);
            cur[d].d = relabel(del, LBL_D);
        }
        memcpy(prev, cur, sizeof prev);
        memcpy(fpm, fcm, sizeof fpm);
        memcpy(fpi, fci, sizeof fpi);
        memcpy(fpd, fcd, sizeof fpd);
    }
    int64_t best = KBIG + 8;
    int bf = 0;
    for (int d = 0; d < BAND; ++d) {
        int64_t k = prev[d].m;
        int f = fpm[d];
        if (prev[d].i < k) { k = prev[d].i; f = fpi[d]; }
        if (prev[d].d < k) { k = prev[d].d; f = fpd[d]; }
        if (k < best) { best = k; bf = f; }
    }
    if (flank) *flank = bf;
    return (int)(best >> 2);
}

static plo_flank_fn g_flank_fn = 0;
void plo_set_flank_fn(plo_flank_fn fn) { g_flank_fn = fn; }

static int do_align(const uint8_t* hap_seg, const uint8_t* read, const uint8_t* qual, int read_len,
                    int ext, int nuc, const uint8_t* open, char* aln1, char* aln2) {
    if (g_align_fn) {
        int firstpos = 0;
        return g_align_fn((const char*)hap_seg, (const char*)read, (const char*)qual, read_len + 15,
                          read_len, ext, nuc, (const char*)open,
                          g_align_traceback ? aln1 : 0, g_align_traceback ? aln2 : 0, &firstpos);
    }
    return plo_band_align(hap_seg, read, qual, read_len, ext, nuc, open);
}

/* One band alignment as calign.pyx:232-237 / 258-263 run it with doCalculateFlankScore = 1:
 * traceback, then the flank cost is taken off a positive score. */
static int do_align_flank(const uint8_t* hap, const uint8_t* gap_open, int hap_len, int hap_flank, int start,
                          const uint8_t* read, const uint8_t* qual, int read_len, int ext, int nuc,
                          char* aln1, char* aln2) {
    int firstpos = 0, s;
    if (g_align_fn)
        s = g_align_fn((const char*)hap + start, (const char*)read, (const char*)qual, read_len + 15, read_len, ext,
                       nuc, (const char*)gap_open + start, aln1, aln2, &firstpos);
    else
        s = plo_band_align_tb(hap + start, read, qual, read_len, ext, nuc, gap_open + start, aln1, aln2, &firstpos);
    if (s > 0) {
        if (g_flank_fn)
            s -= g_flank_fn(hap_len, hap_flank, (const char*)qual, (const char*)gap_open, ext, nuc, firstpos + start,
                            aln1, aln2);
        else
            s -= plo_flank_score(hap_len, hap_flank, qual, gap_open, ext, nuc, firstpos + start, aln1, aln2);
    }
    return s;
}

/* ------------------------------------------------------------------------------------------
 * Homopolymer gap-open table, src/cython/chaplotype.pyx:64-67 evaluated:
 *   per_base_indel_errors -> chr(int(33.5 + 10*log((idx+1)*q)/log(0.1))) - '!'
 * (tests/test_oracle.py recomputes the table from that formula.)
 * -----------------------------------------------------------------------------------------*/
static const uint8_t HOMOPOL_Q[49] = {45, 42, 41, 39, 37, 32, 28, 23, 20, 19, 17, 16, 15, 14, 13, 12, 11,
                                      11, 10, 9,  9,  8,  8,  7,  7,  7,  6,  6,  6,  5,  5,  5,  4,  4,
                                      4,  3,  3,  3,  3,  2,  2,  2,  2,  2,  1,  1,  1,  1,  1};

/* src/cython/chaplotype.pyx:552-590: scan right to left; the run length counts identical
 * bases immediately to the right, saturates at the end of the table, and 'N' never
 * continues a run. */
void plo_gap_open(const uint8_t* hap, int hap_len, uint8_t* out) {
    int homopol = -1, run = 0;
    out[hap_len] = 0;
    for (int i = hap_len - 1; i >= 0; --i) {
        if ((int)hap[i] == homopol) {
            if (run + 1 < 49) run += 1;
        } else {
            run = 0;
        }
        out[i] = HOMOPOL_Q[run];
        homopol = hap[i];
        if (homopol == 'N') homopol = 0;
    }
}

/* ------------------------------------------------------------------------------------------
 * 7-mer hashing, src/cython/calign.pyx:61-90: c = ch & 7; 7 -> 2; two low bits.
 * -----------------------------------------------------------------------------------------*/
static inline uint32_t base_code(uint8_t ch) {
    uint32_t c = ch & 7u;
    if (c == 7u) c = 2u;
    return c & 3u;
}

uint32_t plo_kmer_hash(const uint8_t* seq) {
    uint32_t h = 0;
    for (int i = 0; i < KMER; ++i) h = (h << 2) + base_code(seq[i]);
    return h;
}

typedef struct {
    int16_t* head; /* [HASH_SIZE]: 1 + first position with that hash, 0 = none */
    int16_t* next; /* [hap_len+1]: indexed by 1+position */
    int cap_next;
} HapIndex;

/* src/cython/calign.pyx:94-124: positions 0 .. hap_len-8 are indexed (range(len-7)). */
static void hap_index_build(HapIndex* ix, const uint8_t* hap, int hap_len) {
    memset(ix->head, 0, HASH_SIZE * sizeof(int16_t));
    if (ix->cap_next < hap_len + 1) {
        free(ix->next);
        ix->cap_next = hap_len + 1;
        ix->next = (int16_t*)malloc((size_t)ix->cap_next * sizeof(int16_t));
    }
    memset(ix->next, 0, (size_t)(hap_len + 1) * sizeof(int16_t));
    if (hap_len < KMER) return;
    for (int i = 0; i < hap_len - KMER; ++i) {
        uint32_t h = plo_kmer_hash(hap + i);
        int slot = i + 1;
        if (ix->head[h] == 0) {
            ix->head[h] = (int16_t)slot;
        } else {
            int j = ix->head[h];
            while (ix->next[j] != 0) j = ix->next[j];
            ix->next[j] = (int16_t)slot;
        }
    }
}

/* src/cython/calign.pyx:155-165: hashes of read k-mers 0 .. read_len-8. */
static void read_hashes(const uint8_t* read, int read_len, int16_t* out) {
    uint32_t h = plo_kmer_hash(read);
    out[0] = (int16_t)h;
    for (int i = 1; i < read_len - KMER; ++i) {
        h = ((h << 2) & (HASH_SIZE - 1)) + base_code(read[i + KMER - 1]);
        out[i] = (int16_t)h;
    }
}

/* src/cython/calign.pyx:170-272 with prebuilt index / read hashes (the reference caches
 * both: chaplotype.pyx:637-642).  counts: scratch of >= hap_len+read_len ints. */
static int map_and_align(const uint8_t* read, const uint8_t* qual, const int16_t* rhash, int read_start,
                         int hap_start, int read_len, int hap_len, const uint8_t* hap,
                         const HapIndex* ix, const uint8_t* gap_open, int ext, int nuc, int* counts,
                         char* aln1, char* aln2, int* n_dp, int hap_flank, int do_flank) {
    if (n_dp) *n_dp = 0;
    if (read_len < KMER) return 0; /* :182-183 */
    int maxcount = 0;
    int best_pos = -1;
    int best = PLB_SCORE_NONE;
    /* :196 compares against haplotype-1 and never fires; dropped. */
    memset(counts, 0, (size_t)(hap_len + read_len) * sizeof(int));
    for (int i = 0; i < read_len - KMER; ++i) { /* :209-220 */
        int hidx = ix->head[(uint16_t)rhash[i]];
        while (hidx != 0) {
            int pos = hidx - i - 1;
            int c = ++counts[pos + read_len];
            if (c > maxcount) maxcount = c;
            hidx = ix->next[hidx];
        }
    }
    if (maxcount > 0) { /* :222-247 */
        for (int i = 0; i < hap_len + read_len; ++i) {
            if (counts[i] != maxcount) continue;
            int idx = i - read_len;
            if (idx >= -read_len && idx + read_len + 15 < hap_len) {
                int start = imax(0, idx - 8);
                /* :236-238: the flank cost comes off when doCalculateFlankScore and hapFlank > 0 */
                int s = (do_flank && hap_flank > 0)
                            ? do_align_flank(hap, gap_open, hap_len, hap_flank, start, read, qual, read_len, ext, nuc, aln1, aln2)
                            : do_align(hap + start, read, qual, read_len, ext, nuc, gap_open + start, aln1, aln2);
                if (n_dp) ++*n_dp;
                if (s < best) {
                    best = s;
                    best_pos = idx;
                    if (best == 0) return 0;
                }
            }
        }
    }
    /* :252-267 original mapping position, clamped so the segment ends inside the haplotype */
    int idx0 = imin(read_start - hap_start, hap_len - read_len - 15);
    if (idx0 != best_pos) {
        int start = imax(0, idx0 - 8);
        /* :262-264 tests hapLen > 0 here (not hapFlank); with hapFlank <= 0 the reference has no
         * traceback buffers and would dereference NULL, so callers must pass hap_flank > 0 */
        int s = do_flank ? do_align_flank(hap, gap_open, hap_len, hap_flank, start, read, qual, read_len, ext, nuc, aln1, aln2)
                         : do_align(hap + start, read, qual, read_len, ext, nuc, gap_open + start, aln1, aln2);
        if (n_dp) ++*n_dp;
        if (s < best) best = s;
    }
    return best;
}

/* hash_read (may be NULL = read): the sequence whose 7-mer hashes vote.  In HLA mode the
 * reference clips read/quals/readLen but keeps read.hash of the UNCLIPPED read
 * (chaplotype.pyx:637-638, 647-655), so 7-mer i of the unclipped read votes as if it were
 * 7-mer i of the clipped one. */
int plo_map_and_align_ex(const uint8_t* read, const uint8_t* qual, int read_start, int hap_start,
                         int read_len, int hap_len, const uint8_t* hap, const uint8_t* gap_open,
                         int ext, int nuc, int hap_flank, int do_flank, const uint8_t* hash_read,
                         int hash_read_len, int* n_dp) {
    if (read_len < KMER) {
        if (n_dp) *n_dp = 0;
        return 0;
    }
    if (!hash_read) { hash_read = read; hash_read_len = read_len; }
    HapIndex ix;
    ix.head = (int16_t*)malloc(HASH_SIZE * sizeof(int16_t));
    ix.next = 0;
    ix.cap_next = 0;
    hap_index_build(&ix, hap, hap_len);
    int16_t* rh = (int16_t*)malloc((size_t)(hash_read_len + 1) * sizeof(int16_t));
    read_hashes(hash_read, hash_read_len, rh);
    int* counts = (int*)malloc((size_t)(hap_len + read_len + 1) * sizeof(int));
    char* aln1 = (char*)malloc((size_t)(2 * read_len + 16));
    char* aln2 = (char*)malloc((size_t)(2 * read_len + 16));
    int s = map_and_align(read, qual, rh, read_start, hap_start, read_len, hap_len, hap, &ix, gap_open, ext,
                          nuc, counts, aln1, aln2, n_dp, hap_flank, do_flank);
    free(aln1);
    free(aln2);
    free(counts);
    free(rh);
    free(ix.next);
    free(ix.head);
    return s;
}

int plo_map_and_align(const uint8_t* read, const uint8_t* qual, int read_start, int hap_start,
                      int read_len, int hap_len, const uint8_t* hap, const uint8_t* gap_open,
                      int ext, int nuc, int* n_dp) {
    return plo_map_and_align_ex(read, qual, read_start, hap_start, read_len, hap_len, hap, gap_open, ext, nuc,
                                1, 0, 0, 0, n_dp);
}

/* src/cython/chaplotype.pyx:621-622, 634, 675-676 (useMapQualCap = 0):
 * LL = max(-300, mLTOT*score + log(1 - exp(mLTOT*mapq))). */
double plo_score_to_ll(int score, int mapq) {
    double right = log(1.0 - exp(M_LTOT * (double)mapq));
    double v = M_LTOT * (double)score + right;
    return v > PLB_LL_CAP ? v : PLB_LL_CAP; /* NaN cannot occur: score, mapq finite */
}

/* src/cython/chaplotype.pyx:621-634, 664-676 with useMapQualCap = 1 (--HLATyping): the cap is
 * the log-probability of a wrong mapping, and scores above 100 are flattened smoothly:
 * mLTOT * (99 + (score - 99)^0.5 / 0.5). */
double plo_score_to_ll_hla(int score, int mapq) {
    double cap = M_LTOT * (double)mapq;
    double right = log(1.0 - exp(M_LTOT * (double)mapq));
    const double threshold = 100.0, shape = 0.5;
    double v;
    if ((double)score > threshold)
        v = M_LTOT * (threshold - 1.0 + pow((double)score - threshold + 1.0, shape) / shape);
    else
        v = M_LTOT * (double)score + right;
    return v > cap ? v : cap; /* Python max(cap, v): v only when strictly greater (-inf loses) */
}

/* src/cython/chaplotype.pyx:103-115 */
int plo_overlap(int hap_start, int hap_end, int read_pos, int read_end) {
    int s = imax(hap_start, read_pos);
    int e = imin(hap_end, read_end);
    return e > s ? e - s : -1;
}

/* ------------------------------------------------------------------------------------------
 * S2 over a batch: Haplotype.alignReads, src/cython/chaplotype.pyx:306-377.
 * -----------------------------------------------------------------------------------------*/
static int check_opts(const PlbOptions* opt) {
    if (!opt) return PLB_ERR_ARG;
    if (opt->use_mapq_cap != 0 && opt->use_mapq_cap != 1) return PLB_ERR_ARG;
    if (opt->calc_flank_score != 0 && opt->calc_flank_score != 1) return PLB_ERR_ARG;
    return PLB_OK;
}

int plo_window_loglik(const PlbWindowBatch* b, const PlbOptions* opt, PlbLoglikOut* out, int n_threads,
                      PlbRunStats* stats) {
    int rc = check_opts(opt);
    if (rc) return rc;
    if (!b || !out || !out->ll_off) return PLB_ERR_ARG;
    const int nInd = b->n_individuals;
    int64_t tot_pairs = 0, tot_scored = 0, tot_dp = 0, tot_cells = 0;
    if (n_threads < 1) n_threads = 1;
    int err = 0;
#pragma omp parallel num_threads(n_threads) reduction(+ : tot_pairs, tot_scored, tot_dp, tot_cells)
    {
        HapIndex ix;
        ix.head = (int16_t*)malloc(HASH_SIZE * sizeof(int16_t));
        ix.next = 0;
        ix.cap_next = 0;
        int cap_counts = 0, cap_read = 0, cap_go = 0;
        int64_t cap_rh = 0, cap_slots = 0;
        int64_t* rh_off = 0;
        int* counts = 0;
        int16_t* rh = 0;
        char *aln1 = 0, *aln2 = 0;
        uint8_t* go = 0;
#pragma omp for schedule(dynamic, 16)
        for (int w = 0; w < b->n_windows; ++w) {
            /* read k-mer hashes are computed once per read and window, as the reference caches
             * them in read.hash (chaplotype.pyx:637-638) */
            int64_t ws0 = b->wi_slot_off[(int64_t)w * nInd], ws1 = b->wi_slot_off[(int64_t)(w + 1) * nInd];
            int64_t need = 0;
            int max_rlen = 0;
            for (int64_t s = ws0; s < ws1; ++s) {
                int r = b->slot_read[s];
                int rlen = (int)(b->read_seq_off[r + 1] - b->read_seq_off[r]);
                need += rlen + 1;
                if (rlen > max_rlen) max_rlen = rlen;
            }
            if (cap_rh < need) {
                free(rh);
                cap_rh = need;
                rh = (int16_t*)malloc((size_t)cap_rh * sizeof(int16_t));
            }
            if (cap_slots < ws1 - ws0 + 1) {
                free(rh_off);
                cap_slots = ws1 - ws0 + 1;
                rh_off = (int64_t*)malloc((size_t)cap_slots * sizeof(int64_t));
            }
            if (cap_read < max_rlen + 1) {
                free(aln1); free(aln2);
                cap_read = max_rlen + 1;
                aln1 = (char*)malloc((size_t)(2 * cap_read + 16));
                aln2 = (char*)malloc((size_t)(2 * cap_read + 16));
            }
            {
                int64_t o = 0;
                for (int64_t s = ws0; s < ws1; ++s) {
                    int r = b->slot_read[s];
                    int rlen = (int)(b->read_seq_off[r + 1] - b->read_seq_off[r]);
                    rh_off[s - ws0] = o;
                    if (rlen >= KMER) read_hashes(b->read_seq + b->read_seq_off[r], rlen, rh + o);
                    o += rlen + 1;
                }
            }
            for (int h = b->win_hap_off[w]; h < b->win_hap_off[w + 1]; ++h) {
                const uint8_t* hap = b->hap_seq + b->hap_seq_off[h];
                int hap_len = (int)(b->hap_seq_off[h + 1] - b->hap_seq_off[h]);
                if (hap_len > PLB_MAX_HAP_LEN) { err = PLB_ERR_SHAPE; continue; }
                int hloc = h - b->win_hap_off[w];
                if (cap_go < hap_len + 1) {
                    free(go);
                    cap_go = hap_len + 1;
                    go = (uint8_t*)malloc((size_t)cap_go);
                }
                if (cap_counts < hap_len + max_rlen + 1) {
                    free(counts);
                    cap_counts = hap_len + max_rlen + 1;
                    counts = (int*)malloc((size_t)cap_counts * sizeof(int));
                }
                int built = 0;
                for (int i = 0; i < nInd; ++i) {
                    int wi = w * nInd + i;
                    int64_t s0 = b->wi_slot_off[wi];
                    int T = (int)(b->wi_slot_off[wi + 1] - s0);
                    int n_checked = b->wi_n_good[wi] + b->wi_n_bad[wi]; /* broken mates skip the overlap test */
                    int64_t base = out->ll_off[wi] + (int64_t)hloc * T;
                    for (int t = 0; t < T; ++t) {
                        int r = b->slot_read[s0 + t];
                        int rlen = (int)(b->read_seq_off[r + 1] - b->read_seq_off[r]);
                        ++tot_pairs;
                        int skip = 0;
                        if (t < n_checked) { /* chaplotype.pyx:343-361 */
                            int ov = plo_overlap(b->win_start[w], b->win_end[w], b->read_pos[r], b->read_end[r]);
                            if (b->read_qcfail[r] || ov < KMER) skip = 1;
                        }
                        if (skip) {
                            if (out->ll) out->ll[base + t] = 0.0;
                            if (out->score) out->score[base + t] = -1;
                            continue;
                        }
                        if (!built) { /* lazy, chaplotype.pyx:641-645 */
                            hap_index_build(&ix, hap, hap_len);
                            plo_gap_open(hap, hap_len, go);
                            built = 1;
                        }
                        const uint8_t* rs = b->read_seq + b->read_seq_off[r];
                        const uint8_t* rq = b->read_qual + b->read_seq_off[r];
                        int ndp = 0;
                        int read_start = b->read_pos[r];
                        /* hapFlank = hap.endBufferSize = startPos - hapStart (chaplotype.pyx:604, 609) */
                        int hap_flank = b->win_start[w] - b->hap_start[w];
                        if (opt->use_mapq_cap) { /* chaplotype.pyx:647-655: clip the read to the haplotype */
                            int off1 = b->hap_start[w] - read_start;
                            int off2 = read_start + rlen - b->win_start[w] - hap_len; /* sic: startPos, not hapStart */
                            if (off1 < 0) off1 = 0;
                            if (off2 < 0) off2 = 0;
                            read_start += off1;
                            rlen -= off1 + off2;
                            rs += off1;
                            rq += off1;
                        }
                        if (rlen >= KMER && hap_len < rlen + 15) { err = PLB_ERR_SHAPE; continue; }
                        if (opt->calc_flank_score && hap_flank <= 0) { err = PLB_ERR_ARG; continue; }
                        /* the 7-mer hashes stay those of the unclipped read (read.hash, :637-638) */
                        int sc = map_and_align(rs, rq, rh + rh_off[s0 + t - ws0], read_start, b->hap_start[w], rlen,
                                               hap_len, hap, &ix, go, opt->gap_extend, opt->nuc_prior, counts, aln1,
                                               aln2, &ndp, hap_flank, opt->calc_flank_score);
                        ++tot_scored;
                        tot_dp += ndp;
                        tot_cells += 16 * (int64_t)(rlen > 0 ? rlen : 0);
                        if (out->ll)
                            out->ll[base + t] = opt->use_mapq_cap ? plo_score_to_ll_hla(sc, b->read_mapq[r])
                                                                  : plo_score_to_ll(sc, b->read_mapq[r]);
                        if (out->score) out->score[base + t] = sc;
                    }
                }
            }
        }
        free(go); free(counts); free(rh); free(rh_off); free(aln1); free(aln2);
        free(ix.next); free(ix.head);
    }
    if (stats) {
        stats->n_pairs = tot_pairs;
        stats->n_pairs_scored = tot_scored;
        stats->n_dp = tot_dp;
        stats->cells = tot_cells;
        stats->n_anchor_heavy = stats->n_anchor_verify = stats->n_anchor_exact = 0;
    }
    return err;
}

/* ------------------------------------------------------------------------------------------
 * DiploidGenotype.calculateDataLikelihood, src/cython/cgenotype.pyx:131-189.
 * ll1/ll2: n_total per-read log-likelihoods of the two haplotypes (good|bad|broken).
 * -----------------------------------------------------------------------------------------*/
double plo_genotype_loglik(const double* ll1, const double* ll2, int n_total, int n_good, int homozygous,
                           double* gof, double* hap1_like, double* hap2_like) {
    double likelihood = 0.0, gofsum = 0.0, h1 = 0.0, h2 = 0.0;
    for (int r = 0; r < n_total; ++r) {
        double a = ll1[r], c = ll2[r];
        double la = LOG10E * a, lc = LOG10E * c;
        h1 += la;
        h2 += lc;
        gofsum += la > lc ? la : lc;
        if (homozygous) {
            likelihood += a;                                  /* :164-165 arr1 == arr2 */
        } else if (fabs(a - c) >= 3) {
            likelihood += (LOG_HALF + (a > c ? a : c));       /* :168-169 */
        } else if (fabs(a - c) <= 1e-3) {
            likelihood += a;                                  /* :174-175 */
        } else {
            likelihood += log(0.5 * (exp(a) + exp(c)));       /* :178-179 */
        }
    }
    if (gof) *gof = n_good > 0 ? (-10 * gofsum) / n_good : 0.0; /* :182-185 */
    if (hap1_like) *hap1_like = h1;
    if (hap2_like) *hap2_like = h2;
    return likelihood;
}

/* ------------------------------------------------------------------------------------------
 * Population.call EM loop + EMiteration, src/cython/cpopulation.pyx:678-703, 384-457.
 * gl[i*gl_stride + g]; genotype g <-> (s,r) in the order of cgenotype.pyx:193-218.
 * Returns the number of iterations.
 * -----------------------------------------------------------------------------------------*/
int plo_em(const double* gl, const int32_t* n_reads, int n_ind, int n_hap, int gl_stride, int max_iters,
           double* freq, double* em_post) {
    int G = n_hap * (n_hap + 1) / 2;
    double eps = fmin(1e-3, 1.0 / (n_ind * 2 * 2));
    double max_change = eps + 1;
    double* newf = (double*)malloc((size_t)n_hap * sizeof(double));
    for (int k = 0; k < n_hap; ++k) freq[k] = 1.0 / n_hap;
    int iters = 0;
    while (max_change > eps && iters < max_iters) {
        int n_with = 0;
        for (int i = 0; i < n_ind; ++i) {
            if (n_reads[i] == 0) continue;
            double* csr = em_post + (size_t)i * gl_stride;
            double sum = 0.0;
            ++n_with;
            int g = 0;
            for (int s = 0; s < n_hap; ++s)
                for (int r = s; r < n_hap; ++r, ++g) {
                    double v = gl[(size_t)i * gl_stride + g] * freq[s] * freq[r] * (1 + (r != s));
                    csr[g] = v;
                    sum += v;
                }
            if (sum > 0.0)
                for (g = 0; g < G; ++g) csr[g] /= sum;
        }
        for (int k = 0; k < n_hap; ++k) newf[k] = 0.0;
        for (int i = 0; i < n_ind; ++i) {
            if (n_reads[i] == 0) continue;
            const double* csr = em_post + (size_t)i * gl_stride;
            int g = 0;
            for (int s = 0; s < n_hap; ++s)
                for (int r = s; r < n_hap; ++r, ++g) {
                    newf[s] += csr[g];
                    newf[r] += csr[g];
                }
        }
        max_change = 0.0;
        for (int k = 0; k < n_hap; ++k) {
            newf[k] = newf[k] / (2 * n_with);
            double ch = fabs(freq[k] - newf[k]);
            if (ch > max_change) max_change = ch;
            freq[k] = newf[k];
        }
        ++iters;
    }
    free(newf);
    return iters;
}

/* Population.calculatePosterior, src/cython/cpopulation.pyx:459-594. */
double plo_posterior(const double* gl, const int32_t* n_reads, int n_ind, int n_hap, int gl_stride,
                     const double* freq, const uint64_t* hap_var_mask, int var, double prior) {
    double* fp = (double*)malloc((size_t)n_hap * sizeof(double));
    double sumf = 0.0;
    for (int i = 0; i < n_hap; ++i) {
        if (!((hap_var_mask[i] >> var) & 1u)) {
            fp[i] = freq[i];
            sumf += freq[i];
        } else {
            fp[i] = 0.0;
        }
    }
    if (sumf > 0)
        for (int i = 0; i < n_hap; ++i) fp[i] /= sumf;
    double slv = 0.0, sln = 0.0;
    for (int i = 0; i < n_ind; ++i) {
        if (n_reads[i] == 0) continue;
        double pv = 0.0, pn = 0.0;
        int g = 0;
        for (int r = 0; r < n_hap; ++r)
            for (int s = r; s < n_hap; ++s, ++g) {
                double l = gl[(size_t)i * gl_stride + g];
                double factor = (r != s) ? 2.0 : 1.0;
                pv += (factor * freq[r] * freq[s] * l);
                pn += (factor * fp[r] * fp[s] * l);
            }
        slv += pv > 0 ? log(pv) : -708;
        sln += pn > 0 ? log(pn) : -708;
    }
    free(fp);
    double ratio = fmax(1e-300, exp(sln - slv));
    return round(-10.0 * (log10(ratio * (1.0 - prior)) - log10(prior + ratio * (1.0 - prior))));
}

/* ------------------------------------------------------------------------------------------
 * S3 over a batch: Population.setup + call, src/cython/cpopulation.pyx:197-309, 678-720.
 * -----------------------------------------------------------------------------------------*/
int plo_population_run(const PlbWindowBatch* b, const PlbOptions* opt, PlbPopulationOut* out,
                       PlbLoglikOut* ll_in, int n_threads, PlbRunStats* stats) {
    int rc = check_opts(opt);
    if (rc) return rc;
    if (!b || !out) return PLB_ERR_ARG;
    const int nInd = b->n_individuals;
    const int Hmax = out->max_haps;
    const int Gmax = Hmax * (Hmax + 1) / 2;
    const int64_t nwi = (int64_t)b->n_windows * nInd;

    PlbLoglikOut ll;
    int64_t* off_own = 0;
    double* ll_own = 0;
    if (ll_in && ll_in->ll && ll_in->ll_off) {
        ll = *ll_in;
    } else {
        off_own = (int64_t*)malloc((size_t)(nwi + 1) * sizeof(int64_t));
        int64_t tot = 0;
        for (int w = 0; w < b->n_windows; ++w)
            for (int i = 0; i < nInd; ++i) {
                int64_t wi = (int64_t)w * nInd + i;
                off_own[wi] = tot;
                tot += (int64_t)(b->win_hap_off[w + 1] - b->win_hap_off[w]) *
                       (b->wi_slot_off[wi + 1] - b->wi_slot_off[wi]);
            }
        off_own[nwi] = tot;
        ll_own = (double*)malloc((size_t)(tot > 0 ? tot : 1) * sizeof(double));
        ll.ll_off = off_own;
        ll.ll = ll_own;
        ll.score = ll_in ? ll_in->score : 0;
    }
    rc = plo_window_loglik(b, opt, &ll, n_threads, stats);
    if (rc) { free(off_own); free(ll_own); return rc; }

    if (n_threads < 1) n_threads = 1;
    int err = 0;
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 64)
    for (int w = 0; w < b->n_windows; ++w) {
        int H = b->win_hap_off[w + 1] - b->win_hap_off[w];
        if (H > Hmax || H < 1) { err = PLB_ERR_SHAPE; continue; }
        int G = H * (H + 1) / 2;
        double* gl = (double*)malloc((size_t)nInd * Gmax * sizeof(double));
        double* emp = (double*)calloc((size_t)nInd * Gmax, sizeof(double));
        int32_t* nreads = (int32_t*)malloc((size_t)nInd * sizeof(int32_t));
        double* freq = (double*)calloc((size_t)Hmax, sizeof(double));
        for (int i = 0; i < nInd; ++i) {
            int64_t wi = (int64_t)w * nInd + i;
            int T = (int)(b->wi_slot_off[wi + 1] - b->wi_slot_off[wi]);
            int ngood = b->wi_n_good[wi];
            nreads[i] = ngood;                               /* cpopulation.pyx:286-287 */
            double maxll = -1e7;                             /* :288 */
            const double* L = ll.ll + ll.ll_off[wi];
            int g = 0;
            for (int h1 = 0; h1 < H; ++h1) {
                if (out->hap_like) {
                    double s = 0.0;
                    for (int t = 0; t < T; ++t) s += LOG10E * L[(size_t)h1 * T + t];
                    out->hap_like[((size_t)w * nInd + i) * Hmax + h1] = s;
                }
                for (int h2 = h1; h2 < H; ++h2, ++g) {
                    if (ngood == 0) {                        /* :293-294 */
                        gl[(size_t)i * Gmax + g] = 1.0;
                        if (out->gof) out->gof[((size_t)w * Gmax + g) * nInd + i] = 0.0;
                        continue;
                    }
                    double gof = 0.0;
                    double v = plo_genotype_loglik(L + (size_t)h1 * T, L + (size_t)h2 * T, T, ngood, h1 == h2, &gof,
                                                   0, 0);
                    if (v > maxll) maxll = v;
                    gl[(size_t)i * Gmax + g] = v;
                    if (out->gof) out->gof[((size_t)w * Gmax + g) * nInd + i] = gof;
                }
            }
            for (g = 0; g < G; ++g) {                        /* :304-309 */
                double* p = &gl[(size_t)i * Gmax + g];
                *p = (ngood != 0) ? fmax(1e-300, exp(*p - maxll)) : 1.0;
            }
            for (g = G; g < Gmax; ++g) gl[(size_t)i * Gmax + g] = 0.0;
            if (out->gl_log_max) out->gl_log_max[(size_t)w * nInd + i] = maxll;
        }
        int iters = plo_em(gl, nreads, nInd, H, Gmax, opt->max_em_iters, freq, emp);
        if (out->gl) memcpy(out->gl + (size_t)w * nInd * Gmax, gl, (size_t)nInd * Gmax * sizeof(double));
        if (out->em_post) memcpy(out->em_post + (size_t)w * nInd * Gmax, emp, (size_t)nInd * Gmax * sizeof(double));
        if (out->freq) memcpy(out->freq + (size_t)w * Hmax, freq, (size_t)Hmax * sizeof(double));
        if (out->em_iters) out->em_iters[w] = iters;
        if (out->call) {                                     /* callGenotypes, :623-676 */
            for (int i = 0; i < nInd; ++i) {
                int bestg = -1;
                double bestv = 0.0;
                if (nreads[i] != 0) {
                    const double* src = (opt->use_em_likelihoods == 1 ? emp : gl) + (size_t)i * Gmax;
                    for (int g = 0; g < G; ++g)
                        if (bestg == -1 || src[g] > bestv) { bestv = src[g]; bestg = g; }
                }
                out->call[(size_t)w * nInd + i] = bestg;
            }
        }
        if (out->var_phred && b->max_variants > 0 && b->win_n_var) {
            for (int v = 0; v < b->max_variants; ++v) {
                double ph = 0.0;
                if (v < b->win_n_var[w])
                    ph = plo_posterior(gl, nreads, nInd, H, Gmax, freq, b->hap_var_mask + b->win_hap_off[w], v,
                                       b->var_prior[(size_t)w * b->max_variants + v]);
                out->var_phred[(size_t)w * b->max_variants + v] = ph;
            }
        }
        free(gl); free(emp); free(nreads); free(freq);
    }
    free(off_own);
    free(ll_own);
    return err;
}


/* ------------------------------------------------------------------------------------------
 * N4 — computeGenotypeCallAndLikelihoods, src/cython/vcfutils.pyx:163-334, and what
 * outputCallToVCF derives per sample (vcfutils.pyx:491-548).  "Parity unpinned": restated from the
 * cited lines (the module needs the whole Python-2 object graph).
 * Python semantics kept: max(a, b) returns a unless b > a (so a NaN second argument loses),
 * round() is half away from zero, C division (cdivision=True, src/setup.py:56) may give inf/NaN.
 * -----------------------------------------------------------------------------------------*/
static double py_max(double a, double b) { return b > a ? b : a; }
static double py_min(double a, double b) { return b < a ? b : a; }

int plo_site_genotypes(const PlbWindowBatch* b, const PlbPopulationOut* pop, const PlbSiteBatch* st,
                       PlbSiteOut* out) {
    if (!b || !pop || !st || !out || !pop->gl || !pop->gof || !pop->freq || !b->hap_var_mask) return PLB_ERR_ARG;
    const int nInd = b->n_individuals;
    const int Hm = pop->max_haps, Gm = Hm * (Hm + 1) / 2;
    const int P = out->max_pairs;
    for (int s = 0; s < st->n_sites; ++s) {
        const int w = st->site_win[s];
        const int h0 = b->win_hap_off[w], H = b->win_hap_off[w + 1] - h0;
        const int nV = st->site_var_off[s + 1] - st->site_var_off[s];
        const int32_t* vars = st->site_var + st->site_var_off[s];
        const uint8_t* is_ref = st->hap_is_ref + st->site_hap_off[s];
        if ((nV + 1) * (nV + 2) / 2 > P) return PLB_ERR_SHAPE;
        const double* freq = pop->freq + (size_t)w * Hm;
        for (int i = 0; i < nInd; ++i) {
            const size_t o = (size_t)s * nInd + i;
            const double* gl = pop->gl + ((size_t)w * nInd + i) * Gm;
            if (out->lik) for (int k = 0; k < P; ++k) out->lik[o * P + k] = 0.0;
            if (b->wi_n_good[(size_t)w * nInd + i] == 0) { /* vcfutils.pyx:497-499 */
                if (out->phased) out->phased[o * 2] = out->phased[o * 2 + 1] = -1;
                if (out->post) out->post[o * 3] = out->post[o * 3 + 1] = out->post[o * 3 + 2] = 0.0;
                if (out->phred) out->phred[o * 3] = out->phred[o * 3 + 1] = out->phred[o * 3 + 2] = 0;
                if (out->gof) out->gof[o] = 0.0;
                if (out->gt) out->gt[o * 2] = out->gt[o * 2 + 1] = -1;
                if (out->gl_log10) out->gl_log10[o * 3] = out->gl_log10[o * 3 + 1] = out->gl_log10[o * 3 + 2] = 0.0;
                continue;
            }
            double sum_lik = 0.0, best_gof = 1e6, best_lik = -1.0, nonref = 0.0, ref = 0.0, phased_max = -1e6;
            double liks[3] = {0.0, 0.0, 0.0}, max_lik = 0.0;
            int ph1 = -1, ph2 = -1, pair = 0, have_max = 0;
            for (int i1 = 0; i1 <= nV; ++i1)
                for (int i2 = 0; i2 <= i1; ++i2, ++pair) {
                    double marg = 0.0;
                    int g = 0;
                    for (int a = 0; a < H; ++a)          /* haplotypeIndexes: (a, c), a <= c, cgenotype.pyx:193-218 */
                        for (int c = a; c < H; ++c, ++g) {
                            const int ref1 = is_ref[a], ref2 = is_ref[c];
                            const double factor = (a != c) ? 2.0 : 1.0;
                            int v1h1 = 0, v1h2 = 0, v2h1 = 0, v2h2 = 0, match = 0;
                            if (i1 == 0 && i2 == 0) {
                                match = ref1 && ref2;
                            } else if (i2 == 0) {
                                v1h1 = (int)((b->hap_var_mask[h0 + a] >> vars[i1 - 1]) & 1);
                                v1h2 = (int)((b->hap_var_mask[h0 + c] >> vars[i1 - 1]) & 1);
                                match = (ref2 && v1h1) || (ref1 && v1h2);
                            } else {
                                v1h1 = (int)((b->hap_var_mask[h0 + a] >> vars[i1 - 1]) & 1);
                                v1h2 = (int)((b->hap_var_mask[h0 + c] >> vars[i1 - 1]) & 1);
                                v2h1 = (int)((b->hap_var_mask[h0 + a] >> vars[i2 - 1]) & 1);
                                v2h2 = (int)((b->hap_var_mask[h0 + c] >> vars[i2 - 1]) & 1);
                                match = (v1h1 && v2h2) || (v2h1 && v1h2);
                            }
                            if (!match) continue;
                            const double cur = nInd > 25 ? (factor * freq[a] * freq[c] * gl[g]) : (factor * gl[g]);
                            marg += cur;
                            if (cur > phased_max) { /* :276-316 */
                                phased_max = cur;
                                if (i1 == 0 && i2 == 0) { ph1 = i1; ph2 = i2; }
                                else if (i2 == 0 && i1 != 0) {
                                    if (v1h1) { ph1 = i1; ph2 = i2; }
                                    else if (v1h2) { ph1 = i2; ph2 = i1; }
                                } else if (i2 == i1 && i1 > 0) { ph1 = i1; ph2 = i2; }
                                else if (i2 > 0 && i1 > 0 && i2 != i1) {
                                    if (v1h1 && v2h2) { ph1 = i1; ph2 = i2; }
                                    else if (v1h2 && v2h1) { ph1 = i2; ph2 = i1; }
                                }
                            }
                            const double gf = pop->gof[((size_t)w * Gm + g) * nInd + i];
                            if (gf < best_gof) best_gof = gf;
                        }
                    if (marg > best_lik) best_lik = marg;
                    if ((i1 == 1 && i2 == 0) || (i1 == 1 && i2 == 1)) nonref += marg;
                    else if (i1 == 0 && i2 == 0) ref += marg;
                    sum_lik += marg;
                    if (out->lik) out->lik[o * P + pair] = marg;
                    if (pair < 3) liks[pair] = marg;
                    if (!have_max || marg > max_lik) { max_lik = marg; have_max = 1; }  /* Python max(list) */
                }
            const double gpost = best_lik / sum_lik, npost = nonref / sum_lik, rpost = ref / sum_lik;
            const int q_g = (int)py_min(99, round(-10.0 * log10(py_max(1e-10, 1.0 - gpost))));
            const int q_n = (int)py_min(99, round(-10.0 * log10(py_max(1e-10, 1.0 - npost))));
            const int q_r = (int)py_min(99, round(-10.0 * log10(py_max(1e-10, 1.0 - rpost))));
            int gt1 = ph1, gt2 = ph2;
            double gl3[3] = {-1.0, -1.0, -1.0};
            if (nV == 1) { /* :518-532 */
                if (q_n < st->min_posterior && q_r < st->min_posterior) gt1 = gt2 = -1;
                else if (q_n < st->min_posterior) gt1 = gt2 = 0;
                for (int k = 0; k < 3; ++k) gl3[k] = log10(py_max(liks[k] / max_lik, 1e-300));
            }
            if (out->phased) { out->phased[o * 2] = ph1; out->phased[o * 2 + 1] = ph2; }
            if (out->post) { out->post[o * 3] = gpost; out->post[o * 3 + 1] = npost; out->post[o * 3 + 2] = rpost; }
            if (out->phred) { out->phred[o * 3] = q_g; out->phred[o * 3 + 1] = q_n; out->phred[o * 3 + 2] = q_r; }
            if (out->gof) out->gof[o] = best_gof;
            if (out->gt) { out->gt[o * 2] = gt1; out->gt[o * 2 + 1] = gt2; }
            if (out->gl_log10) for (int k = 0; k < 3; ++k) out->gl_log10[o * 3 + k] = gl3[k];
        }
    }
    return PLB_OK;
}
