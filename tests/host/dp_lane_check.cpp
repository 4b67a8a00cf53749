// CPU check of the packed-lane DP (platypus_b200/csrc/plb_dp.cuh, host emulation of the
// s16x2 intrinsics) against the oracle's cell-by-cell restatement.  Built and run by
// tests/test_host_logic.py; no GPU involved.  Prints "mismatches N".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../platypus_b200/csrc/plb_dp.cuh"
extern "C" int plo_band_align(const uint8_t*, const uint8_t*, const uint8_t*, int, int, int, const uint8_t*);
extern "C" int plo_band_align_tb(const uint8_t*, const uint8_t*, const uint8_t*, int, int, int, const uint8_t*, char*, char*,
                                 int*);
extern "C" int plo_band_align_flank(const uint8_t*, const uint8_t*, const uint8_t*, int, int, int, const uint8_t*, int, int,
                                    int, int*);

static uint64_t rs = 88172645463325252ull;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }

int main(int argc, char** argv) {
    int n_cases = argc > 1 ? atoi(argv[1]) : 3000;
    int bad = 0, bad_gen = 0, n6 = 0, n5 = 0, bad_fl = 0, bad_tb = 0;
    const char* alpha = "ACGT";
    for (int c = 0; c < n_cases; ++c) {
        int L = 9 + rnd() % 260;
        if (c % 7 == 0) L = 9 + rnd() % 12;
        int x0 = rnd() % 20;
        int hapLen = x0 + L + 15 + rnd() % 30;
        std::vector<uint8_t> hap(hapLen), open(hapLen + 1), read(L), qual(L);
        for (auto& b : hap) b = alpha[rnd() % 4];
        if (c % 5 == 0) for (int k = 0; k < 4; ++k) hap[rnd() % hapLen] = 'N';
        const bool big_open = (c % 3 != 0);   // two thirds of the cases: every gap-open >= ext (5-op variant applies)
        for (auto& o : open) o = (uint8_t)(big_open ? 3 + rnd() % 43 : 1 + rnd() % 45);
        open[hapLen] = 0;
        int src = x0 + rnd() % 16, i = src;
        for (int y = 0; y < L; ++y) {
            uint32_t u = rnd() % 1000;
            if (c % 11 == 0) { read[y] = "ACGTN"[rnd() % 5]; continue; }           // unrelated read
            if (u < 20) { read[y] = alpha[rnd() % 4]; ++i; }
            else if (u < 28) { read[y] = alpha[rnd() % 4]; }
            else if (u < 36) { i += 1 + rnd() % 3; read[y] = hap[i < hapLen ? i : hapLen - 1]; ++i; }
            else if (u < 40) { read[y] = 'N'; ++i; }
            else { read[y] = hap[i < hapLen ? i : hapLen - 1]; ++i; }
        }
        for (auto& q : qual) q = (uint8_t)((rnd() % 10 == 0) ? 0 : rnd() % 42);
        int ext = 3, nuc = 2;
        int want = plo_band_align(hap.data() + x0, read.data(), qual.data(), L, ext, nuc, open.data() + x0);
        // device-format staging
        int n = plb::dp_steps(L);
        std::vector<plb::u32> prof(n + 8, 0);
        for (int y = 0; y < L; ++y) prof[y] = plb::make_profile(plb::fast_code(read[y]), qual[y]);
        std::vector<plb::HapRec> rec(hapLen + plb::kRecPad + 64);
        for (size_t x = 0; x < rec.size(); ++x) {
            auto code = [&](size_t p) { return p < (size_t)hapLen ? plb::fast_code(hap[p]) : 4; };
            auto go = [&](size_t p) { return p <= (size_t)hapLen ? (plb::u32)open[p] : 0u; };
            rec[x].gow = go(x) | (go(x + 4) << 16);
            rec[x].sel = plb::make_sel(code(x), code(x + 4));
        }
        int got = plb::band_dp_fast(prof.data(), rec.data() + x0, L, ext, nuc);
        int gen = plb::band_dp_general(hap.data() + x0, open.data() + x0, read.data(), qual.data(), L, ext, nuc);
        if (got != want) { if (++bad < 6) printf("fast mismatch L=%d x0=%d want=%d got=%d\n", L, x0, want, got); }
        // 6-op variant: only defined for segments without 'N'
        bool has_n = false;
        for (int x = x0; x < x0 + L + 15 && x < hapLen; ++x) has_n |= hap[x] == 'N';
        if (!has_n) {
            const int K = 2 * ext + nuc;
            std::vector<plb::u32> prof6(n + 8, 0);
            for (int y = 0; y < L; ++y) prof6[y] = plb::make_profile6(plb::fast_code(read[y]), qual[y], K);
            std::vector<plb::HapRec> rec6(rec.size());
            for (size_t x = 0; x < rec6.size(); ++x) {
                auto code = [&](size_t p) { int c = p < (size_t)hapLen ? plb::fast_code(hap[p]) : 0; return c < 4 ? c : 0; };
                auto go = [&](size_t p) { return (p <= (size_t)hapLen ? (int)open[p] : 0) - ext; };
                rec6[x].gow = plb::pack_s16x2(go(x), go(x + 4));
                rec6[x].sel = plb::make_sel6(code(x), code(x + 4));
            }
            int got6 = plb::band_dp_fast6(prof6.data(), rec6.data() + x0, L, ext, nuc);
            if (got6 != want) { if (++bad < 6) printf("fast6 mismatch L=%d x0=%d want=%d got=%d\n", L, x0, want, got6); }
            ++n6;
            if (big_open) {   // 5-op variant: one three-input min per cell pair
                int got5 = plb::band_dp_fast5(prof6.data(), rec6.data() + x0, L, ext, nuc);
                if (got5 != want) { if (++bad < 6) printf("fast5 mismatch L=%d x0=%d want=%d got=%d\n", L, x0, want, got5); }
                ++n5;
            }
        }
        {   // scope row a2: one-pass flank score and the full traceback against the oracle
            const int flank = 1 + rnd() % (hapLen / 2 + 1);
            int fw = 0, fg = 0;
            const int sw = plo_band_align_flank(hap.data() + x0, read.data(), qual.data(), L, ext, nuc, open.data() + x0, x0,
                                                hapLen, flank, &fw);
            const int sg = plb::band_dp_flank(hap.data() + x0, open.data() + x0, read.data(), qual.data(), L, ext, nuc, x0,
                                              hapLen, flank, &fg);
            if (sw != want || sg != sw || fg != fw) {
                if (++bad_fl < 6) printf("flank mismatch L=%d x0=%d flank=%d want=(%d,%d) got=(%d,%d)\n", L, x0, flank, sw, fw, sg, fg);
            }
            std::vector<char> a1(2 * L + 16), a2(2 * L + 16), b1(2 * L + 16), b2(2 * L + 16);
            std::vector<uint8_t> ptr((size_t)L * 16);
            int fp_w = -1, fp_g = -1;
            const int tw = plo_band_align_tb(hap.data() + x0, read.data(), qual.data(), L, ext, nuc, open.data() + x0, a1.data(),
                                             a2.data(), &fp_w);
            const int tg = plb::band_dp_traceback(hap.data() + x0, open.data() + x0, read.data(), qual.data(), L, ext, nuc,
                                                  ptr.data(), b1.data(), b2.data(), &fp_g);
            if (tw != tg || fp_w != fp_g || strcmp(a1.data(), b1.data()) || strcmp(a2.data(), b2.data())) {
                if (++bad_tb < 6) printf("traceback mismatch L=%d x0=%d\n", L, x0);
            }
        }
        if (gen != want) { if (++bad_gen < 6) printf("general mismatch L=%d want=%d got=%d\n", L, want, gen); }
    }
    printf("mismatches %d general %d of %d (%d through the 6-op variant)\n", bad, bad_gen, n_cases, n6);
    printf("flank mismatches %d traceback mismatches %d (%d cases through the 5-op variant)\n", bad_fl, bad_tb, n5);
    return (bad || bad_gen || bad_fl || bad_tb) ? 1 : 0;
}
