// plb_kmer.cuh — bit-parallel count of the 7-mer votes for ONE offset.
//
// In the reference (src/cython/calign.pyx:206-220) read 7-mer i votes for offset idx exactly when
// the haplotype 7-mer at position i+idx has the same 14-bit hash, i.e. when the seven 2-bit base
// codes (calign.pyx:61-76) at read positions i..i+6 equal those at haplotype positions
// i+idx..i+idx+6.  With both sequences packed 2 bits per base, the number of votes for idx is
//     popcount over i of  AND_{t<7} [ code_r[i+t] == code_h[i+idx+t] ]
// which costs about one instruction per base instead of a table walk per 7-mer.
#pragma once
#include <stdint.h>

#include "plb_dp.cuh"

namespace plb {

#if defined(__CUDA_ARCH__)
PLB_HD u32 fsr(u32 lo, u32 hi, int sh) { return __funnelshift_r(lo, hi, sh); }   // (hi:lo) >> sh, sh in [0,31]
PLB_HD int popc32(u32 x) { return __popc(x); }
#else
PLB_HD u32 fsr(u32 lo, u32 hi, int sh) { return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
PLB_HD int popc32(u32 x) { return __builtin_popcount(x); }
#endif

// 2-bit base code of the reference's hash (calign.pyx:69-74): c = ch & 7; 7 -> 2; c & 3.
PLB_HD u32 kmer_base_code(uint8_t ch) {
    u32 c = ch & 7u;
    if (c == 7u) c = 2u;
    return c & 3u;
}

constexpr int kPackPadWords = 3;  // zero words before and after every packed sequence

// rpk : packed read  (base i at bits 2*(i&15) of word i>>4), readable for words [0, ceil(L/16)+2]
// hpk : packed haplotype, same layout, readable for words [-kPackPadWords, ceil(hapLen/16)+kPackPadWords)
// nk_read = readLen-7 and nk_hap = hapLen-7 are the numbers of indexed 7-mers (calign.pyx:109,164)
// Returns #{ i in [0,nk_read) : 0 <= i+idx < nk_hap and 7-mer i of the read == 7-mer i+idx of the haplotype }.
PLB_HD int count_offset_bits(const u32* __restrict__ rpk, const u32* __restrict__ hpk, int nk_read, int nk_hap, int idx) {
    const int lo = idx < 0 ? -idx : 0;
    int hi = nk_hap - idx;
    if (hi > nk_read) hi = nk_read;
    if (hi <= lo) return 0;
    const int w0 = lo >> 4, w1 = (hi - 1) >> 4;  // read words that hold valid 7-mer starts
    // per-base match bits (bit 2j set iff base 16w+j of the read equals base 16w+j+idx of the haplotype)
    auto match_word = [&](int w) -> u32 {
        const int hb = 16 * w + idx;              // first haplotype base under this read word (>= -15)
        const int hw = hb >> 4;                   // arithmetic shift: floor
        const int sh = 2 * (hb & 15);
        const u32 h = fsr(hpk[hw], hpk[hw + 1], sh);
        const u32 e = rpk[w] ^ h;
        return ~(e | (e >> 1)) & 0x55555555u;
    };
    int c = 0;
    const u32 mlo = 0xFFFFFFFFu << (2 * (lo - 16 * w0));            // lo - 16 w0 in [0, 15]
    const u32 mhi = 0xFFFFFFFFu >> (2 * (16 * w1 + 16 - hi));        // 16 w1 + 16 - hi in [0, 15]
    u32 m0 = match_word(w0), m1 = match_word(w0 + 1);
    u32 a0 = m0 & fsr(m0, m1, 2);                 // bases i, i+1
    for (int w = w0; w <= w1; ++w) {
        const u32 m2 = match_word(w + 2);
        const u32 a1 = m1 & fsr(m1, m2, 2);
        const u32 b = a0 & fsr(a0, a1, 4);        // i .. i+3
        const u32 cc = b & fsr(a0, a1, 8);        // i .. i+5
        u32 d = cc & fsr(m0, m1, 12);             // i .. i+6
        // keep 7-mer starts in [lo, hi): only the first and the last word can hold starts outside it
        if (w == w0) d &= mlo;
        if (w == w1) d &= mhi;
        c += popc32(d);
        m0 = m1;
        m1 = m2;
        a0 = a1;
    }
    return c;
}

}  // namespace plb
