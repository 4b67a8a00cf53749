// plb_dp.cuh — banded affine-gap min-plus alignment, one alignment per thread.
//
// Replaces the forward pass of fastAlignmentRoutine (reference: src/c/align.c:77-521).
// The reference sweeps anti-diagonals with 8 SSE2 int16 lanes; here one CUDA thread owns
// one (read, haplotype segment) alignment and keeps the same 2x8 anti-diagonal lanes in
// registers as packed s16x2 words, updated with the sm_100a packed-halfword min/add
// instructions (VIADD.16x2, VIMNMX.S16x2, VIADDMNMX.S16x2) and PRMT.
//
// Lane layout.  Step pair t computes two anti-diagonals:
//   even vector lane i : cell (x = t+i,   y = t-i)   diagonal d = 2i
//   odd  vector lane i : cell (x = t+1+i, y = t-i)   diagonal d = 2i+1
// Register k of a vector packs lanes (k, k+4) as (lo16, hi16).  With that packing a
// one-lane shift of a vector is a register renaming plus ONE PRMT.
//
// Substitution cost without compares.  Each read row y has a 32-bit "profile"
//   prof[y] = cost vs haplotype base A | C<<8 | G<<16 | T<<24   (0 where it matches read[y])
// and each haplotype position x a PRMT selector that picks byte code(x) for the lo lane and
// byte code(x+4) for the hi lane (or a zero byte when the haplotype base is 'N', which the
// reference scores as a free match, align.c:175-178).  One PRMT yields two cells' costs.
//
// Everything here is in phred units (the reference works in phred*4 with the traceback
// label in the low bits; traceback is not computed here and never changes the score).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PLB_HD __host__ __device__ __forceinline__
#else
#define PLB_HD static inline
#ifndef __align__
#define __align__(n) __attribute__((aligned(n)))
#endif
#endif

namespace plb {

typedef uint32_t u32;

constexpr u32 kInf2 = 0x70007000u;  // "+inf" in both lanes; INF + any one-step cost < 0x8000
constexpr int kInf = 0x7000;
constexpr int kScoreBig = 1 << 28;

// ---- packed s16x2 primitives (host emulation is used by the CPU unit test of the lane logic)
#if defined(__CUDA_ARCH__)
PLB_HD u32 vadd2(u32 a, u32 b) { return __vadd2(a, b); }
PLB_HD u32 vmin2(u32 a, u32 b) { return __vmins2(a, b); }
PLB_HD u32 vaddmin2(u32 a, u32 b, u32 c) { return __viaddmin_s16x2(a, b, c); }  // min(a+b, c)
// PTX prmt, default mode: selector nibble bit 3 replicates the sign bit of the selected byte
// (used to produce zero bytes); __byte_perm() masks that bit off, so go through inline PTX.
PLB_HD u32 prmt(u32 a, u32 b, u32 s) {
    u32 d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}
#else
PLB_HD u32 vadd2(u32 a, u32 b) {
    return ((a + b) & 0xFFFFu) | ((((a >> 16) + (b >> 16)) & 0xFFFFu) << 16);
}
PLB_HD u32 vmin2(u32 a, u32 b) {
    int16_t al = (int16_t)(a & 0xFFFF), bl = (int16_t)(b & 0xFFFF);
    int16_t ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
    return (u32)(uint16_t)(al < bl ? al : bl) | ((u32)(uint16_t)(ah < bh ? ah : bh) << 16);
}
PLB_HD u32 vaddmin2(u32 a, u32 b, u32 c) { return vmin2(vadd2(a, b), c); }
PLB_HD u32 prmt(u32 a, u32 b, u32 s) {
    uint64_t v = ((uint64_t)b << 32) | a;
    u32 r = 0;
    for (int i = 0; i < 4; ++i) {
        u32 n = (s >> (4 * i)) & 0xF;
        u32 byte = (u32)((v >> (8 * (n & 7))) & 0xFF);
        if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
#endif

// Haplotype record for position x (8 bytes, built when a window is staged):
//   .x = gap-open pair  open[x] | open[x+4] << 16            (reference: localgapopen, align.c:185-187)
//   .y = PRMT selector  code(x) | 0x80 | (4+code(x+4)) << 8 | 0x8000, nibble 8 where the base is 'N'
struct __align__(8) HapRec {
    u32 gow;
    u32 sel;
};

PLB_HD u32 hap_sel_nibble_lo(int code) { return code < 4 ? (u32)code : 8u; }
PLB_HD u32 hap_sel_nibble_hi(int code) { return code < 4 ? (u32)(4 + code) : 8u; }
PLB_HD u32 make_sel(int code_lo, int code_hi) {
    return hap_sel_nibble_lo(code_lo) | 0x80u | (hap_sel_nibble_hi(code_hi) << 8) | 0x8000u;
}
// read base + quality -> profile word.  code 0..3 = A,C,G,T; anything else mismatches all.
PLB_HD u32 make_profile(int code, u32 qual) {
    u32 p = qual * 0x01010101u;
    if (code < 4) p &= ~(0xFFu << (8 * code));
    return p;
}
// exact byte -> code for the fast path alphabet; haplotype 'N' is the wildcard (4), every
// other byte is 5 (never matches anything; windows whose haplotypes contain such bytes take
// the general path, reads may contain them freely).
PLB_HD int fast_code(uint8_t ch) {
    // branch-free: (ch>>1)&3 maps A,C,T,G to 0,1,2,3; the byte table confirms the exact letter
    const u32 idx = (ch >> 1) & 3u;
    const u32 e = (0x47544341u >> (8 * idx)) & 0xFFu;   // 'A','C','T','G'
    const u32 c = (0x02030100u >> (8 * idx)) & 0xFFu;   //  0 , 1 , 3 , 2
    return e == ch ? (int)c : (ch == 'N' ? 4 : 5);
}

// number of step pairs executed for a read of length L (multiple of 8, >= 24)
PLB_HD int dp_steps(int L) {
    int n = (L + 8 + 7) & ~7;
    return n < 24 ? 24 : n;
}
constexpr int kMinFastLen = 9;   // shorter reads take the general path
constexpr int kRecPad = 48;      // zero records appended after each haplotype

struct DpState {
    u32 ME[4], IE[4], DE[4], MIE[4];
    u32 MO[4], IO[4], DO[4], MIO[4];
    u32 P[8];                // profile ring, slot = row & 7
    u32 Wg[8], Wp[8], Ws[8]; // record ring (gap-open, gap-open+nucprior, selector), slot = index & 7
    u32 acc[4];              // running min of the last-row cells
    u32 NM[4];               // 0 in the lane whose row is L-1, 0x7FFF elsewhere
};

// One group of 8 step pairs starting at t0 (multiple of 8).
// FIRST: t0 == 0, applies the y = -1 boundary (free start on all 16 diagonals, align.c:244-250).
// LAST : collects min over the band cells of row L-1 (align.c:261-288, 416-443).
template <bool FIRST, bool LAST>
PLB_HD void dp_group(DpState& s, const u32* __restrict__ prof, const HapRec* __restrict__ rec, int t0, int L,
                     u32 ext2, u32 extp2, u32 nuc2) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = t0 + j;
        // new profile row t and record t+4 enter the rings
        s.P[j & 7] = prof[t];
        {
            HapRec r = rec[t + 4];
            s.Wg[(j + 4) & 7] = r.gow;
            s.Wp[(j + 4) & 7] = vadd2(r.gow, nuc2);
            s.Ws[(j + 4) & 7] = r.sel;
        }
        if (LAST) {
            if (t == L - 1) s.NM[0] &= 0xFFFF0000u;
        }
        // ---------------- even half: cells (t+i, t-i) ----------------
        u32 Dt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 pa = s.P[(j - k) & 7], pb = s.P[(j - k - 4) & 7];
            const int w0 = (j + k) & 7, w1 = (j + k + 1) & 7;
            u32 B = vmin2(s.MIE[k], s.DE[k]);                                  // B(x-1,y-1)
            u32 sub = prmt(pa, pb, s.Ws[w0]);
            u32 Mn = vadd2(B, sub);
            u32 In = vaddmin2(s.IO[k], extp2, vadd2(s.MO[k], s.Wp[w0]));       // from (x, y-1), no D->I
            Dt[k] = vaddmin2(s.DO[k], ext2, vadd2(s.MIO[k], s.Wg[w1]));        // from (x-1, y), I->D allowed
            s.ME[k] = Mn;
            s.IE[k] = In;
            s.MIE[k] = vmin2(Mn, In);
        }
        s.DE[3] = Dt[2];
        s.DE[2] = Dt[1];
        s.DE[1] = Dt[0];
        s.DE[0] = prmt(Dt[3], kInf2, 0x1054);  // lane 0 <- +inf, lane 4 <- lane 3
        if (FIRST && j < 7) {                  // lane j+1 sits on row y = -1: B := 0, M,I := inf
            const int k = (j + 1) & 3;
            const u32 keep = (j + 1) < 4 ? 0xFFFF0000u : 0x0000FFFFu;
            const u32 inf1 = (j + 1) < 4 ? 0x00007000u : 0x70000000u;
            s.ME[k] = (s.ME[k] & keep) | inf1;
            s.IE[k] = (s.IE[k] & keep) | inf1;
            s.MIE[k] = (s.MIE[k] & keep) | inf1;
            s.DE[k] = (s.DE[k] & keep);
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s.acc[k] = vmin2(s.acc[k], vmin2(s.MIE[k], s.DE[k]) | s.NM[k]);
        }
        // ---------------- odd half: cells (t+1+i, t-i) ----------------
        u32 It[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 pa = s.P[(j - k) & 7], pb = s.P[(j - k - 4) & 7];
            const int w0 = (j + k) & 7, w1 = (j + k + 1) & 7;
            u32 B = vmin2(s.MIO[k], s.DO[k]);
            u32 sub = prmt(pa, pb, s.Ws[w1]);
            u32 Mn = vadd2(B, sub);
            u32 Dn = vaddmin2(s.DE[k], ext2, vadd2(s.MIE[k], s.Wg[w1]));
            It[k] = vaddmin2(s.IE[k], extp2, vadd2(s.ME[k], s.Wp[w0]));
            s.MO[k] = Mn;
            s.DO[k] = Dn;
        }
        s.IO[0] = It[1];
        s.IO[1] = It[2];
        s.IO[2] = It[3];
        s.IO[3] = prmt(It[0], kInf2, 0x7632);  // lane 3 <- lane 4, lane 7 <- +inf
#pragma unroll
        for (int k = 0; k < 4; ++k) s.MIO[k] = vmin2(s.MO[k], s.IO[k]);
        if (FIRST && j < 7) {                  // row y = -1, even x: M := 0 (align.c:244-250), I := inf
            const int k = (j + 1) & 3;
            const u32 keep = (j + 1) < 4 ? 0xFFFF0000u : 0x0000FFFFu;
            const u32 inf1 = (j + 1) < 4 ? 0x00007000u : 0x70000000u;
            s.MO[k] = (s.MO[k] & keep);
            s.IO[k] = (s.IO[k] & keep) | inf1;
            s.MIO[k] = (s.MIO[k] & keep) | inf1;
            s.DO[k] = (s.DO[k] & keep);
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) s.acc[k] = vmin2(s.acc[k], vmin2(s.MIO[k], s.DO[k]) | s.NM[k]);
            // hot lane moves up by one for the next step pair
            u32 n3 = s.NM[3];
            s.NM[3] = s.NM[2];
            s.NM[2] = s.NM[1];
            s.NM[1] = s.NM[0];
            s.NM[0] = prmt(n3, 0x7FFF7FFFu, 0x1054);
        }
    }
}

// prof : profile words of the read, rows 0..L-1, zero-padded to dp_steps(L) rows
// rec  : haplotype records, rec[0] is segment position x = 0; must be readable up to
//        index dp_steps(L)+4 (zero padding past the haplotype end)
// Requires L >= kMinFastLen.  Returns the alignment score in phred units.
PLB_HD int band_dp_fast(const u32* __restrict__ prof, const HapRec* __restrict__ rec, int L, int ext, int nuc) {
    DpState s;
    const u32 ext2 = (u32)ext * 0x00010001u;
    const u32 nuc2 = (u32)nuc * 0x00010001u;
    const u32 extp2 = (u32)(ext + nuc) * 0x00010001u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.ME[k] = s.IE[k] = s.DE[k] = s.MIE[k] = kInf2;
        s.MO[k] = s.IO[k] = s.DO[k] = s.MIO[k] = kInf2;
        s.acc[k] = 0x7FFF7FFFu;
        s.NM[k] = 0x7FFF7FFFu;
    }
    // "step -1": lane 0 of both vectors sits on row y = -1
    s.DE[0] = 0x70000000u;   // even lane 0: D := 0 so that B = 0, M = I = inf
    s.MO[0] = 0x70000000u;   // odd lane 0 (x = 0, even): M := 0
    s.DO[0] = 0x70000000u;
#pragma unroll
    for (int m = 0; m < 8; ++m) s.P[m] = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        HapRec r = rec[m];
        s.Wg[m] = r.gow;
        s.Wp[m] = vadd2(r.gow, nuc2);
        s.Ws[m] = r.sel;
    }
#pragma unroll
    for (int m = 4; m < 8; ++m) s.Wg[m] = s.Wp[m] = s.Ws[m] = 0;

    const int n = dp_steps(L);
    dp_group<true, false>(s, prof, rec, 0, L, ext2, extp2, nuc2);
    int t0 = 8;
    for (; t0 < n - 16; t0 += 8) dp_group<false, false>(s, prof, rec, t0, L, ext2, extp2, nuc2);
    dp_group<false, true>(s, prof, rec, t0, L, ext2, extp2, nuc2);
    dp_group<false, true>(s, prof, rec, t0 + 8, L, ext2, extp2, nuc2);
    u32 a = vmin2(vmin2(s.acc[0], s.acc[1]), vmin2(s.acc[2], s.acc[3]));
    int lo = (int)(a & 0xFFFF), hi = (int)(a >> 16);
    return lo < hi ? lo : hi;
}

// ---------------------------------------------------------------------------------------------
// 6-op variant (haplotype groups without 'N').
//
// Every state is stored relative to the frame  f(x,y) = (ext+nuc)*y + ext*x :  S~ = S - f.
// An insertion step (y+1) adds ext+nuc and a deletion step (x+1) adds ext, so in this frame
// both gap-extension additions vanish and both gap-open terms become the same per-position
// constant open[x] - ext:
//     M~ = B~' + (sub - K)                 K = 2*ext + nuc (the diagonal step crosses both axes)
//     I~ = min(I~', M~' + (open[x]-ext))   one VIADDMNMX
//     D~ = min(D~', MI~' + (open[x]-ext))  one VIADDMNMX
//     MI~ = min(M~, I~),  B~ = min(MI~, D~)
// i.e. 6 packed ops per 2 cells instead of 8.  The profile bytes hold sub-K as SIGNED bytes and
// the PRMT selector sign-extends them (nibble 8|c), which leaves no spare byte for the 'N'
// wildcard - groups with an 'N' use the 8-op variant above.
// Absolute score = relative + f; the last-row extraction adds ext*d per diagonal and the common
// (2*ext+nuc)*(L-1) at the end.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PLB_HD u32 vmax2(u32 a, u32 b) { return __vmaxs2(a, b); }
#else
PLB_HD u32 vmax2(u32 a, u32 b) {
    int16_t al = (int16_t)(a & 0xFFFF), bl = (int16_t)(b & 0xFFFF);
    int16_t ah = (int16_t)(a >> 16), bh = (int16_t)(b >> 16);
    return (u32)(uint16_t)(al > bl ? al : bl) | ((u32)(uint16_t)(ah > bh ? ah : bh) << 16);
}
#endif

constexpr int kMaxFastLen = 2000;  // relative values reach -(2*ext+nuc)*L: stay far inside int16

PLB_HD u32 make_sel6(int code_lo, int code_hi) {  // codes 0..3 only
    const u32 lo = (u32)code_lo, hi = (u32)(4 + code_hi);
    return lo | ((8u | lo) << 4) | (hi << 8) | ((8u | hi) << 12);
}
// profile word of signed bytes: mismatch q-K, match -K.  code > 3 (N or any other byte) mismatches all.
PLB_HD u32 make_profile6(int code, int qual, int K) {
    const u32 mis = (u32)(uint8_t)(int8_t)(qual - K), mat = (u32)(uint8_t)(int8_t)(-K);
    u32 p = mis * 0x01010101u;
    if (code < 4) p = (p & ~(0xFFu << (8 * code))) | (mat << (8 * code));
    return p;
}
PLB_HD u32 pack_s16x2(int lo, int hi) { return ((u32)lo & 0xFFFFu) | ((u32)hi << 16); }

struct DpState6 {
    u32 ME[4], IE[4], DE[4], MIE[4];
    u32 MO[4], IO[4], DO[4], MIO[4];
    u32 P[8];
    u32 Wg[8], Ws[8];
    u32 acc[4];
    u32 NM[4];   // 0x8000 in the lane whose row is L-1, 0x7FFF elsewhere
};

template <bool FIRST, bool LAST>
PLB_HD void dp_group6(DpState6& s, const u32* __restrict__ prof, const HapRec* __restrict__ rec, int t0, int L,
                      int ext, int extp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = t0 + j;
        s.P[j & 7] = prof[t];
        {
            HapRec r = rec[t + 4];
            s.Wg[(j + 4) & 7] = r.gow;
            s.Ws[(j + 4) & 7] = r.sel;
        }
        if (LAST) {
            if (t == L - 1) s.NM[0] = (s.NM[0] & 0xFFFF0000u) | 0x8000u;
        }
        // ---------------- even half: cells (t+i, t-i) ----------------
        u32 Dt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 pa = s.P[(j - k) & 7], pb = s.P[(j - k - 4) & 7];
            const int w0 = (j + k) & 7, w1 = (j + k + 1) & 7;
            u32 B = vmin2(s.MIE[k], s.DE[k]);
            u32 Mn = vadd2(B, prmt(pa, pb, s.Ws[w0]));
            u32 In = vaddmin2(s.MO[k], s.Wg[w0], s.IO[k]);
            Dt[k] = vaddmin2(s.MIO[k], s.Wg[w1], s.DO[k]);
            s.ME[k] = Mn;
            s.IE[k] = In;
            s.MIE[k] = vmin2(Mn, In);
        }
        s.DE[3] = Dt[2];
        s.DE[2] = Dt[1];
        s.DE[1] = Dt[0];
        s.DE[0] = prmt(Dt[3], kInf2, 0x1054);
        if (FIRST && j < 7) {  // lane j+1 on row y = -1, x = 2j+1: B := 0 (absolute), M, I := inf
            const int k = (j + 1) & 3;
            const bool lo = (j + 1) < 4;
            const u32 keep = lo ? 0xFFFF0000u : 0x0000FFFFu;
            const u32 inf1 = lo ? 0x00007000u : 0x70000000u;
            const u32 zero = ((u32)(extp - ext * (2 * j + 1)) & 0xFFFFu) << (lo ? 0 : 16);  // 0 - f(x,-1)
            s.ME[k] = (s.ME[k] & keep) | inf1;
            s.IE[k] = (s.IE[k] & keep) | inf1;
            s.MIE[k] = (s.MIE[k] & keep) | inf1;
            s.DE[k] = (s.DE[k] & keep) | zero;
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 dc = pack_s16x2(ext * 2 * k, ext * 2 * (k + 4));   // + ext*d, d = 2*lane
                s.acc[k] = vmin2(s.acc[k], vmax2(vadd2(vmin2(s.MIE[k], s.DE[k]), dc), s.NM[k]));
            }
        }
        // ---------------- odd half: cells (t+1+i, t-i) ----------------
        u32 It[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 pa = s.P[(j - k) & 7], pb = s.P[(j - k - 4) & 7];
            const int w0 = (j + k) & 7, w1 = (j + k + 1) & 7;
            u32 B = vmin2(s.MIO[k], s.DO[k]);
            u32 Mn = vadd2(B, prmt(pa, pb, s.Ws[w1]));
            u32 Dn = vaddmin2(s.MIE[k], s.Wg[w1], s.DE[k]);
            It[k] = vaddmin2(s.ME[k], s.Wg[w0], s.IE[k]);
            s.MO[k] = Mn;
            s.DO[k] = Dn;
        }
        s.IO[0] = It[1];
        s.IO[1] = It[2];
        s.IO[2] = It[3];
        s.IO[3] = prmt(It[0], kInf2, 0x7632);
#pragma unroll
        for (int k = 0; k < 4; ++k) s.MIO[k] = vmin2(s.MO[k], s.IO[k]);
        if (FIRST && j < 7) {  // row y = -1, x = 2j+2 (even): M := 0 (absolute), I := inf
            const int k = (j + 1) & 3;
            const bool lo = (j + 1) < 4;
            const u32 keep = lo ? 0xFFFF0000u : 0x0000FFFFu;
            const u32 inf1 = lo ? 0x00007000u : 0x70000000u;
            const u32 zero = ((u32)(extp - ext * (2 * j + 2)) & 0xFFFFu) << (lo ? 0 : 16);
            s.MO[k] = (s.MO[k] & keep) | zero;
            s.IO[k] = (s.IO[k] & keep) | inf1;
            s.MIO[k] = (s.MIO[k] & keep) | inf1;
            s.DO[k] = (s.DO[k] & keep) | zero;
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 dc = pack_s16x2(ext * (2 * k + 1), ext * (2 * (k + 4) + 1));  // d = 2*lane + 1
                s.acc[k] = vmin2(s.acc[k], vmax2(vadd2(vmin2(s.MIO[k], s.DO[k]), dc), s.NM[k]));
            }
            u32 n3 = s.NM[3];
            s.NM[3] = s.NM[2];
            s.NM[2] = s.NM[1];
            s.NM[1] = s.NM[0];
            s.NM[0] = prmt(n3, 0x7FFF7FFFu, 0x1054);
        }
    }
}

// prof : make_profile6 words (K = 2*ext+nuc), rows 0..L-1, zero-padded to dp_steps(L) rows
// rec  : records with .gow = pack(open[x]-ext, open[x+4]-ext), .sel = make_sel6(code(x), code(x+4))
// Requires kMinFastLen <= L <= kMaxFastLen and no 'N' in the segment.
PLB_HD int band_dp_fast6(const u32* __restrict__ prof, const HapRec* __restrict__ rec, int L, int ext, int nuc) {
    DpState6 s;
    const int extp = ext + nuc;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.ME[k] = s.IE[k] = s.DE[k] = s.MIE[k] = kInf2;
        s.MO[k] = s.IO[k] = s.DO[k] = s.MIO[k] = kInf2;
        s.acc[k] = 0x7FFF7FFFu;
        s.NM[k] = 0x7FFF7FFFu;
    }
    // "step -1": lane 0 of both vectors sits on row y = -1 (even lane: x = -1, odd lane: x = 0)
    s.DE[0] = 0x70000000u | ((u32)(extp + ext) & 0xFFFFu);   // absolute 0 at (x=-1,y=-1)
    s.MO[0] = 0x70000000u | ((u32)extp & 0xFFFFu);           // absolute 0 at (x=0,y=-1)
    s.DO[0] = s.MO[0];
#pragma unroll
    for (int m = 0; m < 8; ++m) s.P[m] = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        HapRec r = rec[m];
        s.Wg[m] = r.gow;
        s.Ws[m] = r.sel;
    }
#pragma unroll
    for (int m = 4; m < 8; ++m) s.Wg[m] = s.Ws[m] = 0;

    const int n = dp_steps(L);
    dp_group6<true, false>(s, prof, rec, 0, L, ext, extp);
    int t0 = 8;
    for (; t0 < n - 16; t0 += 8) dp_group6<false, false>(s, prof, rec, t0, L, ext, extp);
    dp_group6<false, true>(s, prof, rec, t0, L, ext, extp);
    dp_group6<false, true>(s, prof, rec, t0 + 8, L, ext, extp);
    u32 a = vmin2(vmin2(s.acc[0], s.acc[1]), vmin2(s.acc[2], s.acc[3]));
    int lo = (int)(int16_t)(a & 0xFFFF), hi = (int)(int16_t)(a >> 16);
    // relative -> absolute: + (ext+nuc)*(L-1) + ext*x with x = (L-1) + d; ext*d was added per lane
    return (lo < hi ? lo : hi) + (extp + ext) * (L - 1);
}

// ---------------------------------------------------------------------------------------------
// 5-op variant: ONE three-input min per cell pair instead of two two-input ones.
//
// The reference opens a deletion from min(M, I) of the cell to the left (align.c:320-329), so the 6-op form
// keeps MI = min(M, I) next to B = min(MI, D).  Where every gap-open penalty of the window is >= the
// gap-extension penalty, opening from B gives the same value:
//     min(D + ext, min(M, I, D) + open) = min(D + ext, MI + open, D + open) = min(D + ext, MI + open)
// because D + open >= D + ext.  Then only B is needed - for the diagonal step and for the deletion - and it
// is one VIMNMX3 of the three fresh states.  Per cell pair: PRMT, VIADD, VIADDMNMX, VIADDMNMX, VIMNMX3
// = 4 ALU-pipe ops + 1 add (the 6-op form has 5 + 1).  The homopolymer table (chaplotype.pyx:64-67) drops
// below ext = 3 only inside runs of 40 or more identical bases; k_prep flags such windows and they keep
// the 6-op form.  Same profile / record format as band_dp_fast6.
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PLB_HD u32 vmin3(u32 a, u32 b, u32 c) { return __vimin3_s16x2(a, b, c); }
#else
PLB_HD u32 vmin3(u32 a, u32 b, u32 c) { return vmin2(vmin2(a, b), c); }
#endif

struct DpState5 {
    u32 ME[4], IE[4], DE[4], BE[4];
    u32 MO[4], IO[4], DO[4], BO[4];
    u32 P[8];
    u32 Wg[8], Ws[8];
    u32 acc[4];
    u32 NM[4];   // 0x8000 in the lane whose row is L-1, 0x7FFF elsewhere
};

template <bool FIRST, bool LAST>
PLB_HD void dp_group5(DpState5& s, const u32* __restrict__ prof, const HapRec* __restrict__ rec, int t0, int L,
                      int ext, int extp) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int t = t0 + j;
        s.P[j & 7] = prof[t];
        {
            HapRec r = rec[t + 4];
            s.Wg[(j + 4) & 7] = r.gow;
            s.Ws[(j + 4) & 7] = r.sel;
        }
        if (LAST) {
            if (t == L - 1) s.NM[0] = (s.NM[0] & 0xFFFF0000u) | 0x8000u;
        }
        // ---------------- even half: cells (t+i, t-i) ----------------
        u32 Dt[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 pa = s.P[(j - k) & 7], pb = s.P[(j - k - 4) & 7];
            const int w0 = (j + k) & 7, w1 = (j + k + 1) & 7;
            const u32 Mn = vadd2(s.BE[k], prmt(pa, pb, s.Ws[w0]));
            const u32 In = vaddmin2(s.MO[k], s.Wg[w0], s.IO[k]);
            Dt[k] = vaddmin2(s.BO[k], s.Wg[w1], s.DO[k]);
            s.ME[k] = Mn;
            s.IE[k] = In;
        }
        s.DE[3] = Dt[2];
        s.DE[2] = Dt[1];
        s.DE[1] = Dt[0];
        s.DE[0] = prmt(Dt[3], kInf2, 0x1054);
#pragma unroll
        for (int k = 0; k < 4; ++k) s.BE[k] = vmin3(s.ME[k], s.IE[k], s.DE[k]);
        if (FIRST && j < 7) {  // lane j+1 on row y = -1, x = 2j+1: B := 0 (absolute), M, I := inf
            const int k = (j + 1) & 3;
            const bool lo = (j + 1) < 4;
            const u32 keep = lo ? 0xFFFF0000u : 0x0000FFFFu;
            const u32 inf1 = lo ? 0x00007000u : 0x70000000u;
            const u32 zero = ((u32)(extp - ext * (2 * j + 1)) & 0xFFFFu) << (lo ? 0 : 16);  // 0 - f(x,-1)
            s.ME[k] = (s.ME[k] & keep) | inf1;
            s.IE[k] = (s.IE[k] & keep) | inf1;
            s.DE[k] = (s.DE[k] & keep) | inf1;
            s.BE[k] = (s.BE[k] & keep) | zero;
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 dc = pack_s16x2(ext * 2 * k, ext * 2 * (k + 4));   // + ext*d, d = 2*lane
                s.acc[k] = vmin2(s.acc[k], vmax2(vadd2(s.BE[k], dc), s.NM[k]));
            }
        }
        // ---------------- odd half: cells (t+1+i, t-i) ----------------
        u32 It[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const u32 pa = s.P[(j - k) & 7], pb = s.P[(j - k - 4) & 7];
            const int w0 = (j + k) & 7, w1 = (j + k + 1) & 7;
            const u32 Mn = vadd2(s.BO[k], prmt(pa, pb, s.Ws[w1]));
            const u32 Dn = vaddmin2(s.BE[k], s.Wg[w1], s.DE[k]);
            It[k] = vaddmin2(s.ME[k], s.Wg[w0], s.IE[k]);
            s.MO[k] = Mn;
            s.DO[k] = Dn;
        }
        s.IO[0] = It[1];
        s.IO[1] = It[2];
        s.IO[2] = It[3];
        s.IO[3] = prmt(It[0], kInf2, 0x7632);
#pragma unroll
        for (int k = 0; k < 4; ++k) s.BO[k] = vmin3(s.MO[k], s.IO[k], s.DO[k]);
        if (FIRST && j < 7) {  // row y = -1, x = 2j+2 (even): M := 0 and B := 0 (absolute), I, D := inf
            const int k = (j + 1) & 3;
            const bool lo = (j + 1) < 4;
            const u32 keep = lo ? 0xFFFF0000u : 0x0000FFFFu;
            const u32 inf1 = lo ? 0x00007000u : 0x70000000u;
            const u32 zero = ((u32)(extp - ext * (2 * j + 2)) & 0xFFFFu) << (lo ? 0 : 16);
            s.MO[k] = (s.MO[k] & keep) | zero;
            s.IO[k] = (s.IO[k] & keep) | inf1;
            s.DO[k] = (s.DO[k] & keep) | inf1;
            s.BO[k] = (s.BO[k] & keep) | zero;
        }
        if (LAST) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const u32 dc = pack_s16x2(ext * (2 * k + 1), ext * (2 * (k + 4) + 1));  // d = 2*lane + 1
                s.acc[k] = vmin2(s.acc[k], vmax2(vadd2(s.BO[k], dc), s.NM[k]));
            }
            u32 n3 = s.NM[3];
            s.NM[3] = s.NM[2];
            s.NM[2] = s.NM[1];
            s.NM[1] = s.NM[0];
            s.NM[0] = prmt(n3, 0x7FFF7FFFu, 0x1054);
        }
    }
}

// Requires kMinFastLen <= L <= kMaxFastLen, no 'N' in the segment and open[x] >= ext for every x of it.
PLB_HD int band_dp_fast5(const u32* __restrict__ prof, const HapRec* __restrict__ rec, int L, int ext, int nuc) {
    DpState5 s;
    const int extp = ext + nuc;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        s.ME[k] = s.IE[k] = s.DE[k] = s.BE[k] = kInf2;
        s.MO[k] = s.IO[k] = s.DO[k] = s.BO[k] = kInf2;
        s.acc[k] = 0x7FFF7FFFu;
        s.NM[k] = 0x7FFF7FFFu;
    }
    // "step -1": lane 0 of both vectors sits on row y = -1 (even lane: x = -1, odd lane: x = 0)
    s.BE[0] = 0x70000000u | ((u32)(extp + ext) & 0xFFFFu);   // absolute 0 at (x=-1,y=-1)
    s.MO[0] = 0x70000000u | ((u32)extp & 0xFFFFu);           // absolute 0 at (x=0,y=-1)
    s.BO[0] = s.MO[0];
#pragma unroll
    for (int m = 0; m < 8; ++m) s.P[m] = 0;
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        HapRec r = rec[m];
        s.Wg[m] = r.gow;
        s.Ws[m] = r.sel;
    }
#pragma unroll
    for (int m = 4; m < 8; ++m) s.Wg[m] = s.Ws[m] = 0;

    const int n = dp_steps(L);
    dp_group5<true, false>(s, prof, rec, 0, L, ext, extp);
    int t0 = 8;
    for (; t0 < n - 16; t0 += 8) dp_group5<false, false>(s, prof, rec, t0, L, ext, extp);
    dp_group5<false, true>(s, prof, rec, t0, L, ext, extp);
    dp_group5<false, true>(s, prof, rec, t0 + 8, L, ext, extp);
    u32 a = vmin2(vmin2(s.acc[0], s.acc[1]), vmin2(s.acc[2], s.acc[3]));
    int lo = (int)(int16_t)(a & 0xFFFF), hi = (int)(int16_t)(a >> 16);
    return (lo < hi ? lo : hi) + (extp + ext) * (L - 1);
}

// General path: arbitrary bytes, any read length >= 1.  Same recurrence cell by cell
// (SURVEY §3.3), 32-bit scalars.  hap/open point at segment position 0.
PLB_HD int band_dp_general(const uint8_t* __restrict__ hap, const uint8_t* __restrict__ open,
                           const uint8_t* __restrict__ read, const uint8_t* __restrict__ qual, int L, int ext,
                           int nuc) {
    int Mp[16], Ip[16], Dp[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) Mp[d] = Ip[d] = Dp[d] = kScoreBig;
    for (int y = 0; y < L; ++y) {
        const int rb = read[y], q = qual[y];
        int Mc[16], Ic[16], Dc[16];
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const int x = y + d;
            const int hb = hap[x], go = open[x];
            const int sub = (hb == 'N' || hb == rb) ? 0 : q;
            int diag = Mp[d] < Ip[d] ? Mp[d] : Ip[d];
            diag = diag < Dp[d] ? diag : Dp[d];
            if (y == 0) diag = 0;
            int m = diag + sub;
            int ins;
            if (y == 0) {
                ins = (x & 1) ? kScoreBig : go + nuc;
            } else if (d < 15) {
                int a = Ip[d + 1] + ext, b = Mp[d + 1] + go;
                ins = (a < b ? a : b) + nuc;
            } else {
                ins = kScoreBig;
            }
            int del;
            if (d >= 1) {
                int mi = Mc[d - 1] < Ic[d - 1] ? Mc[d - 1] : Ic[d - 1];
                int a = Dc[d - 1] + ext, b = mi + go;
                del = a < b ? a : b;
            } else {
                del = kScoreBig;
            }
            Mc[d] = m < kScoreBig ? m : kScoreBig;
            Ic[d] = ins < kScoreBig ? ins : kScoreBig;
            Dc[d] = del < kScoreBig ? del : kScoreBig;
        }
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            Mp[d] = Mc[d];
            Ip[d] = Ic[d];
            Dp[d] = Dc[d];
        }
    }
    int best = kScoreBig;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        int b = Mp[d] < Ip[d] ? Mp[d] : Ip[d];
        b = b < Dp[d] ? b : Dp[d];
        best = best < b ? best : b;
    }
    return best;
}


// ---------------------------------------------------------------------------------------------
// Scope row a2 / N2: traceback and calculateFlankScore (src/c/align.c:344-365, 493-644).
//
// With traceback on, every value of the reference carries the label of its own state in its two
// low bits (M=0, I=1, D=3; scores are x4) and every `min` compares value and label, so equal
// scores resolve to the smaller label; the label surviving a cell's min is that state's
// back-pointer.  key = 4*score + label reproduces that order exactly.
//
// band_dp_flank computes the flank score WITHOUT storing back-pointers or walking them: each
// state carries, next to its key, the in-flank cost of the path its back-pointers describe
// (a min over keys selects exactly the predecessor the traceback would follow, and
// calculateFlankScore charges each alignment column the same cost the recurrence charged,
// counted only where the column's haplotype coordinate lies in a flank).  One 32-bit word per
// state: key in bits 14.., flank cost in bits 0..13 (flank <= score < 15872).  Candidates of one
// min never share a key (their labels differ), so the packed min is the min by key.
//   column coordinate: M and D at cell (x, y) -> x;  I -> x + 1 (align.c:621-631)
// Returns the score; *flank_out = calculateFlankScore of the traceback alignment.
// ---------------------------------------------------------------------------------------------
constexpr u32 kFlInf = 0xFC000000u;
constexpr u32 kFlLblMask = 3u << 14;

PLB_HD u32 fl_min(u32 a, u32 b) { return a < b ? a : b; }
PLB_HD u32 fl_fix(u32 v, u32 lbl) { v = v < kFlInf ? v : kFlInf; return (v & ~kFlLblMask) | (lbl << 14); }

PLB_HD int band_dp_flank(const uint8_t* __restrict__ hap, const uint8_t* __restrict__ open,
                         const uint8_t* __restrict__ read, const uint8_t* __restrict__ qual, int L, int ext, int nuc,
                         int start, int hap_len, int hap_flank, int* flank_out) {
    u32 M[16], I[16], D[16];
#pragma unroll
    for (int d = 0; d < 16; ++d) M[d] = I[d] = D[d] = kFlInf;
    const int hi_lo = hap_len - hap_flank;   // columns >= hi_lo or < hap_flank are flank columns
    // fm bit d = column (start + y + d) lies in a flank, for d = 0..16
    u32 fm = 0;
    for (int d = 0; d <= 16; ++d) {
        const int hx = start + d;
        fm |= (u32)((hx < hap_flank) | (hx >= hi_lo)) << d;
    }
    for (int y = 0; y < L; ++y) {
        const u32 rb = read[y], q = qual[y];
        u32 mp = kFlInf, ip = kFlInf, dp = kFlInf;   // states of cell (x-1, y), i.e. diagonal d-1 of this row
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const int x = y + d;
            const u32 hb = hap[x], go = open[x];
            const u32 wx = 65536u + ((fm >> d) & 1u);        // cost weight: 4<<14 for the key + 1 if in flank
            const u32 wx1 = 65536u + ((fm >> (d + 1)) & 1u);
            const u32 sub = (hb == 'N' || hb == rb) ? 0u : q;
            u32 diag = fl_min(M[d], fl_min(I[d], D[d]));
            if (y == 0) diag = 0u;
            const u32 m = diag + sub * wx;
            u32 ins;
            if (y == 0) {
                ins = (x & 1) ? kFlInf : (go + (u32)nuc) * wx1;
            } else if (d < 15) {
                ins = fl_min(I[d + 1] + (u32)ext * wx1, M[d + 1] + go * wx1) + (u32)nuc * wx1;
            } else {
                ins = kFlInf;
            }
            u32 del;
            if (d >= 1) {
                del = fl_min(dp + (u32)ext * wx, fl_min(mp, ip) + go * wx);
            } else {
                del = kFlInf;
            }
            mp = fl_fix(m, 0u);
            ip = fl_fix(ins, 1u);
            dp = fl_fix(del, 3u);
            M[d] = mp;
            I[d] = ip;
            D[d] = dp;
        }
        const int hx = start + y + 17;
        fm = (fm >> 1) | ((u32)((hx < hap_flank) | (hx >= hi_lo)) << 16);
    }
    u32 best = 0xFFFFFFFFu;
#pragma unroll
    for (int d = 0; d < 16; ++d) {   // first diagonal with the smallest labelled value, align.c:261-288, 416-443
        const u32 k = fl_min(M[d], fl_min(I[d], D[d]));
        if ((k >> 14) < (best >> 14)) best = k;
    }
    if (flank_out) *flank_out = (int)(best & 0x3FFFu);
    return (int)(best >> 16);
}

// Score of one band alignment as mapAndAlignReadToHaplotype uses it with doCalculateFlankScore = 1
// (calign.pyx:232-238, 258-264): a positive score loses its flank part.
PLB_HD int band_dp_flank_adjusted(const uint8_t* __restrict__ hap, const uint8_t* __restrict__ open,
                                  const uint8_t* __restrict__ read, const uint8_t* __restrict__ qual, int L, int ext,
                                  int nuc, int start, int hap_len, int hap_flank) {
    int fl = 0;
    const int s = band_dp_flank(hap + start, open + start, read, qual, L, ext, nuc, start, hap_len, hap_flank, &fl);
    return s > 0 ? s - fl : s;
}

// Full traceback (plb_fast_align with aln1/aln2): forward pass storing one byte of back-pointers
// per cell (bits 0-1 M, 2-3 I, 6-7 D, the packing of align.c:346-348) in ptr[L*16], then the walk
// of align.c:523-577.  aln1/aln2 must hold 2L+16 bytes.  Returns the score.
PLB_HD int band_dp_traceback(const uint8_t* __restrict__ hap, const uint8_t* __restrict__ open,
                             const uint8_t* __restrict__ read, const uint8_t* __restrict__ qual, int L, int ext,
                             int nuc, uint8_t* __restrict__ ptr, char* __restrict__ aln1, char* __restrict__ aln2,
                             int* firstpos) {
    int M[16], I[16], D[16];   // keys 4*score + label
    const int big = kScoreBig;
#pragma unroll
    for (int d = 0; d < 16; ++d) M[d] = I[d] = D[d] = big;
    for (int y = 0; y < L; ++y) {
        const int rb = read[y], q = qual[y];
        int mp = big, ip = big, dp = big;
#pragma unroll
        for (int d = 0; d < 16; ++d) {
            const int x = y + d;
            const int hb = hap[x], go = open[x];
            const int sub = (hb == 'N' || hb == rb) ? 0 : q;
            int diag = M[d] < I[d] ? M[d] : I[d];
            diag = diag < D[d] ? diag : D[d];
            if (y == 0) diag = 0;
            int m = diag + 4 * sub;
            int ins;
            if (y == 0) {
                ins = (x & 1) ? big : 4 * (go + nuc);
            } else if (d < 15) {
                const int a = I[d + 1] + 4 * ext, b = M[d + 1] + 4 * go;
                ins = (a < b ? a : b) + 4 * nuc;
            } else {
                ins = big;
            }
            int del;
            if (d >= 1) {
                const int mi = mp < ip ? mp : ip;
                const int a = dp + 4 * ext, b = mi + 4 * go;
                del = a < b ? a : b;
            } else {
                del = big;
            }
            ptr[y * 16 + d] = (uint8_t)((m & 3) | ((ins & 3) << 2) | ((del & 3) << 6));
            mp = m < big ? (m & ~3) : big;
            ip = ins < big ? ((ins & ~3) | 1) : big + 1;
            dp = del < big ? ((del & ~3) | 3) : big + 3;
            M[d] = mp;
            I[d] = ip;
            D[d] = dp;
        }
    }
    int best = big + 8, bd = 0;
#pragma unroll
    for (int d = 0; d < 16; ++d) {
        int k = M[d] < I[d] ? M[d] : I[d];
        k = k < D[d] ? k : D[d];
        if (k < best) {
            best = k;
            bd = d;
        }
    }
    int state = best & 3;
    int cx = L - 1 + bd, cy = L - 1, n = 0;
    while (cy >= 0) {
        const int d = cx - cy;
        const int p = (d >= 0 && d < 16) ? ptr[cy * 16 + d] : 0;
        const int ns = (p >> (2 * state)) & 3;
        if (state == 0) {
            aln1[n] = (char)hap[cx];
            aln2[n] = (char)read[cy];
            --cx;
            --cy;
        } else if (state == 1) {
            aln1[n] = '-';
            aln2[n] = (char)read[cy];
            --cy;
        } else {
            aln1[n] = (char)hap[cx];
            aln2[n] = '-';
            --cx;
        }
        state = ns;
        ++n;
    }
    aln1[n] = 0;
    aln2[n] = 0;
    if (firstpos) *firstpos = cx + 1;
    for (int i = 0, j = n - 1; i < j; ++i, --j) {
        char t = aln1[i];
        aln1[i] = aln1[j];
        aln1[j] = t;
        t = aln2[i];
        aln2[i] = aln2[j];
        aln2[j] = t;
    }
    return best >> 2;
}

}  // namespace plb
