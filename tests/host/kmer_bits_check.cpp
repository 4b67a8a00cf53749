// CPU check of the bit-parallel per-offset vote count (platypus_b200/csrc/plb_kmer.cuh) against a
// direct restatement of the reference's vote rule (calign.pyx:206-220).  Prints "mismatches N".
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../platypus_b200/csrc/plb_kmer.cuh"

static uint64_t rs = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17; return (uint32_t)(rs >> 11); }
static uint32_t hash7(const uint8_t* p) { uint32_t h = 0; for (int i = 0; i < 7; ++i) h = (h << 2) + plb::kmer_base_code(p[i]); return h; }

static std::vector<plb::u32> pack(const std::vector<uint8_t>& s, int pad) {
    std::vector<plb::u32> w((s.size() + 15) / 16 + 2 * pad, 0);
    for (size_t i = 0; i < s.size(); ++i) w[pad + (i >> 4)] |= plb::kmer_base_code(s[i]) << (2 * (i & 15));
    return w;
}

int main(int argc, char** argv) {
    int n_cases = argc > 1 ? atoi(argv[1]) : 2000, bad = 0;
    const char* alpha = "ACGTNacgtRY";
    for (int c = 0; c < n_cases; ++c) {
        int L = 7 + rnd() % 300, H = 7 + rnd() % 600;
        std::vector<uint8_t> hap(H), read(L);
        int na = (c % 3 == 0) ? 2 : (c % 3 == 1 ? 4 : 11);   // low-complexity cases give many matches
        for (auto& b : hap) b = alpha[rnd() % na];
        int src = rnd() % H;
        for (int i = 0; i < L; ++i) read[i] = (rnd() % 20 == 0 || src + i >= H) ? alpha[rnd() % na] : hap[src + i];
        auto rp = pack(read, plb::kPackPadWords), hp = pack(hap, plb::kPackPadWords);
        int nkr = L - 7, nkh = H - 7;
        for (int t = 0; t < 12; ++t) {
            int idx = (t < 4) ? src + (int)(rnd() % 7) - 3 : (int)(rnd() % (H + L + 40)) - L - 20;
            int want = 0;
            for (int i = 0; i < nkr; ++i) {
                int p = i + idx;
                if (p >= 0 && p < nkh && hash7(&read[i]) == hash7(&hap[p])) ++want;
            }
            int got = plb::count_offset_bits(rp.data() + plb::kPackPadWords, hp.data() + plb::kPackPadWords, nkr, nkh, idx);
            if (got != want && ++bad < 8) printf("mismatch L=%d H=%d idx=%d want=%d got=%d\n", L, H, idx, want, got);
        }
    }
    printf("mismatches %d of %d\n", bad, n_cases * 12);
    return bad ? 1 : 0;
}
