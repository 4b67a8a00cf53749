// plb_synth.cuh — "synth-v1d": the synthetic windows of SURVEY §8d generated ON THE DEVICE, in place, into a resident batch.
//
// Measurement support for BASELINE config 5 (≈30 M windows x 2000 samples: the inputs of even a 1/1000 subsample are
// ~1 TB and cannot come over PCIe; SURVEY §8d asks for generation on the device).  Same recipe as the host generator
// (platypus_b200/synth.py "synth-v1") with a counter-based hash as random source, so any window can be regenerated on
// any GPU from (seed, window id) alone:
//   reference segment of hapLen+16 iid ACGT, 15 % chance of one homopolymer run of 4-12 in the middle third;
//   haplotype 0 = reference; haplotype h >= 1 = reference with its own SNP in the central 50 bp and, w.p. 0.3, a 1-3 bp
//   insertion or deletion three bases further on (positions distinct per haplotype, so no two haplotypes are equal);
//   per individual a genotype (g1, g2) uniform over haplotype pairs; per read: source g1 or g2, start uniform in
//   [0, hapLen-L-16], qualities 90 % U[25,40] / 10 % U[2,24], substitutions w.p. 10^(-q/10), 0.1 % + 0.1 % per base 1-bp
//   insertion / deletion errors, mapq 85 % 60 / 10 % U[20,59] / 5 % U[0,19], pos = hapStart + start + jitter (0 w.p. 0.9,
//   else U[-5,5]).
// The kernels only WRITE inputs; what the likelihood path computes from them is checked like any other input (bench.py
// downloads sampled windows with plb_batch_download and runs the CPU oracle on them).
#pragma once

namespace plb {

__device__ __forceinline__ uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// uniform 32 bits for (seed, window, stream, counter)
__device__ __forceinline__ uint32_t rnd32(uint64_t seed, uint64_t win, uint32_t stream, uint32_t ctr) {
    return (uint32_t)(mix64(mix64(seed ^ (win * 0x9E3779B97F4A7C15ull)) ^ ((uint64_t)stream << 32 | ctr)) >> 32);
}

__device__ __forceinline__ uint64_t rnd64(uint64_t seed, uint64_t win, uint32_t stream, uint32_t ctr) {
    return mix64(mix64(seed ^ (win * 0x9E3779B97F4A7C15ull)) ^ ((uint64_t)stream << 32 | ctr));
}

__constant__ uint32_t c_sub_thresh[48];   // P(substitution | quality q) * 2^32, q = 0..47

constexpr int kSynthCentre = 50;

// One block per window: reference segment, haplotypes, window coordinates, variant masks / priors.
__global__ void __launch_bounds__(128) k_synth_windows(DevBatch b, uint64_t seed, int64_t first_window) {
    const int w = blockIdx.x;
    const uint64_t wg = (uint64_t)(first_window + w);
    const int h0 = b.win_hap_off[w], H = b.win_hap_off[w + 1] - h0;
    const int hl = (int)(b.hap_seq_off[h0 + 1] - b.hap_seq_off[h0]);
    extern __shared__ uint8_t s_ref[];   // hl + 16
    const int tid = threadIdx.x;
    const char* acgt = "ACGT";
    for (int x = tid; x < hl + 16; x += blockDim.x) s_ref[x] = (uint8_t)acgt[rnd32(seed, wg, 0, (uint32_t)x) & 3];
    __syncthreads();
    if (tid == 0) {
        const uint32_t u = rnd32(seed, wg, 1, 0);
        if (u < (uint32_t)(0.15 * 4294967296.0)) {
            const int run = 4 + (int)(rnd32(seed, wg, 1, 1) % 9u);
            const int third = hl / 3;
            const int p = third + (int)(rnd32(seed, wg, 1, 2) % (uint32_t)max(1, third - run));
            for (int k = 1; k < run; ++k) s_ref[p + k] = s_ref[p];
        }
    }
    __syncthreads();
    const int c0 = (hl - kSynthCentre) / 2;
    const int a = (int)(rnd32(seed, wg, 1, 3) % (uint32_t)kSynthCentre);
    const int hs = 100000 + (int)(wg % 2000000ull) * 1000;
    if (tid == 0) {
        ((int32_t*)b.hap_start)[w] = hs;
        ((int32_t*)b.win_start)[w] = hs + c0;
        ((int32_t*)b.win_end)[w] = hs + c0 + kSynthCentre;
    }
    int n_var = 0;   // identical in every thread (the draws are deterministic)
    for (int g = 0; g < H; ++g) {
        uint8_t* out = (uint8_t*)b.hap_seq + b.hap_seq_off[h0 + g];
        uint64_t mask = 0;
        int p1 = -1, p2 = -1, kind2 = 0, n2 = 0;   // own SNP at p1; optional indel anchored at p2 (kind2 1 = ins, 2 = del)
        uint8_t alt = 0;
        if (g > 0) {
            p1 = c0 + (a + 7 * g) % kSynthCentre;
            const int code = s_ref[p1] == 'A' ? 0 : s_ref[p1] == 'C' ? 1 : s_ref[p1] == 'G' ? 2 : 3;
            alt = (uint8_t)acgt[(code + 1 + (int)(rnd32(seed, wg, 2, (uint32_t)g) % 3u)) & 3];
            mask |= 1ull << n_var;
            if (b.var_prior && tid == 0) ((double*)b.var_prior)[(size_t)w * b.max_variants + n_var] = 1e-3;
            ++n_var;
            if (rnd32(seed, wg, 3, (uint32_t)g) < (uint32_t)(0.3 * 4294967296.0)) {
                p2 = c0 + (a + 7 * g + 3) % kSynthCentre;
                kind2 = 1 + (int)(rnd32(seed, wg, 4, (uint32_t)g) & 1u);
                n2 = 1 + (int)(rnd32(seed, wg, 5, (uint32_t)g) % 3u);
                mask |= 1ull << n_var;
                if (b.var_prior && tid == 0) ((double*)b.var_prior)[(size_t)w * b.max_variants + n_var] = 1e-4;
                ++n_var;
            }
        }
        if (b.hap_var_mask && tid == 0) ((uint64_t*)b.hap_var_mask)[h0 + g] = mask;
        // output position y -> source: walk is cheap enough to do per thread in closed form
        for (int y = tid; y < hl; y += blockDim.x) {
            uint8_t c;
            if (kind2 == 1 && y > p2 && y <= p2 + n2) {            // inserted bases follow the anchor p2
                c = (uint8_t)acgt[rnd32(seed, wg, 6 + (uint32_t)g, (uint32_t)(y - p2)) & 3];
            } else {
                int x = y;
                if (kind2 == 1 && y > p2 + n2) x = y - n2;          // after an insertion
                if (kind2 == 2 && y > p2) x = y + n2;               // after a deletion of n2 bases behind the anchor
                c = s_ref[x];
                if (x == p1) c = alt;
            }
            out[y] = c;
        }
    }
    if (b.win_n_var && tid == 0) ((int32_t*)b.win_n_var)[w] = n_var;
    if (b.var_prior)
        for (int v = n_var + tid; v < b.max_variants; v += blockDim.x) ((double*)b.var_prior)[(size_t)w * b.max_variants + v] = 0.0;
}

// One WARP per read (slot s of the batch = read s of the pool): lanes along the read, 32 bases per round, so that the
// stores coalesce.  Indel errors make a base's source position depend on the events before it; that is an exclusive
// prefix sum of (deletion ? 2 : insertion ? 0 : 1) over the earlier bases - a warp scan with a carry between rounds.
__global__ void __launch_bounds__(128) k_synth_reads(DevBatch b, uint64_t seed, int64_t first_window) {
    const int64_t s = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (s >= b.n_slots) return;   // warp-uniform
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xFFFFFFFFu;
    const int wi = b.slot_wi[s];
    const int nInd = b.n_individuals;
    const int w = wi / nInd, ind = wi % nInd;
    const int t = (int)(s - b.wi_slot_off[wi]);
    const uint64_t wg = (uint64_t)(first_window + w);
    const int h0 = b.win_hap_off[w], H = b.win_hap_off[w + 1] - h0;
    const int hl = (int)(b.hap_seq_off[h0 + 1] - b.hap_seq_off[h0]);
    const int64_t ro = b.read_seq_off[s];
    const int L = (int)(b.read_seq_off[s + 1] - ro);
    const uint32_t stream = 1024u + (uint32_t)ind;        // per individual (window-level draws use streams < 1024)
    const uint32_t base_ctr = (uint32_t)t << 12;          // per read of the individual (L <= 4095 draws per kind)
    const int g1 = (int)(rnd32(seed, wg, stream, 0xFFFFFFF0u) % (uint32_t)H);
    const int g2 = (int)(rnd32(seed, wg, stream, 0xFFFFFFF1u) % (uint32_t)H);
    const uint32_t r0 = rnd32(seed, wg, stream, base_ctr);
    const int src = (r0 & 1u) ? g1 : g2;
    const int span = hl - L - 16 + 1;
    const int idx = span > 0 ? (int)(rnd32(seed, wg, stream, base_ctr + 1) % (uint32_t)span) : 0;
    const uint8_t* hap = b.hap_seq + b.hap_seq_off[h0 + src];
    uint8_t* rs = (uint8_t*)b.read_seq + ro;
    uint8_t* rq = (uint8_t*)b.read_qual + ro;
    const char* acgt = "ACGT";
    int carry = idx;   // source position of the round's first base, before its own deletion
    for (int k0 = 0; k0 < L; k0 += 32) {
        const int k = k0 + lane;
        // two hashes per base: 64 bits for the event (e) and the substitution test / random base (v), 32 for the quality
        const uint64_t hv = rnd64(seed, wg, stream + (2u << 20), base_ctr + (uint32_t)k);
        const uint32_t e = (uint32_t)(hv >> 32), v = (uint32_t)hv;
        const uint32_t u = rnd32(seed, wg, stream + (1u << 20), base_ctr + (uint32_t)k);   // quality
        const int q = (u % 10u) ? 25 + (int)((u >> 8) % 16u) : 2 + (int)((u >> 8) % 23u);
        const bool ins = k < L && e < 4294967u;                        // 0.1 %
        const bool del = k < L && !ins && e < 2u * 4294967u;           // 0.1 %
        const int step = k < L ? (del ? 2 : ins ? 0 : 1) : 0;          // source bases this base consumes
        int incl = step;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(FULL, incl, o);
            if (lane >= o) incl += n;
        }
        const int x = carry + incl - step + (del ? 1 : 0);             // a deletion skips one source base first
        carry += __shfl_sync(FULL, incl, 31);
        if (k < L) {
            uint8_t c = ins ? (uint8_t)acgt[v & 3] : hap[min(x, hl - 1)];
            if ((v >> 2 << 2) < c_sub_thresh[q]) {                     // substitution: one of the three other bases
                const int code = c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : 3;
                c = (uint8_t)acgt[(code + 1 + (int)((e >> 12) % 3u)) & 3];
            }
            rs[k] = c;
            rq[k] = (uint8_t)q;
        }
    }
    if (lane == 0) {
        const uint32_t m = rnd32(seed, wg, stream, base_ctr + 2);
        const uint32_t mm = m % 100u;
        const int mapq = mm < 85u ? 60 : mm < 95u ? 20 + (int)((m >> 8) % 40u) : (int)((m >> 8) % 20u);
        const uint32_t j = rnd32(seed, wg, stream, base_ctr + 3);
        const int jit = (j % 10u) ? 0 : -5 + (int)((j >> 8) % 11u);
        const int hs = 100000 + (int)(wg % 2000000ull) * 1000;
        ((int32_t*)b.read_pos)[s] = hs + idx + jit;
        ((int32_t*)b.read_end)[s] = hs + idx + jit + L;
        ((uint8_t*)b.read_mapq)[s] = (uint8_t)mapq;
        ((uint8_t*)b.read_qcfail)[s] = 0;
    }
}

}  // namespace plb

extern "C" int plb_synth_fill_device(PlbContext* c, PlbDeviceBatch* db, uint64_t seed, int64_t first_window) {
    if (!c || !db) return set_err(PLB_ERR_ARG, "NULL argument");
    int rc = require_idle(c, "plb_synth_fill_device");
    if (rc) return rc;
    const DevBatch& d = db->d;
    if (d.n_windows == 0) return PLB_OK;
    if (db->packed || db->qual_bits || db->shares_reads) return set_err(PLB_ERR_UNSUPPORTED, "plb_synth_fill_device needs an ASCII batch that owns its reads");
    if (d.n_slots != d.n_reads) return set_err(PLB_ERR_ARG, "plb_synth_fill_device: slot s must be read s (one slot per pool read)");
    if (db->have_var && d.max_variants < 2 * (db->max_haps - 1))
        return set_err(PLB_ERR_SHAPE, "plb_synth_fill_device: max_variants %d < 2 * (haplotypes - 1)", d.max_variants);
    if (db->max_hap_len + 16 > 40000) return set_err(PLB_ERR_SHAPE, "haplotypes too long for the generator");
    CU(cudaSetDevice(c->device));
    static bool table_ready[64] = {false};
    if (!table_ready[c->device & 63]) {
        uint32_t th[48];
        for (int q = 0; q < 48; ++q) {
            const double p = std::pow(10.0, -q / 10.0);
            th[q] = p >= 1.0 ? 0xFFFFFFFCu : (uint32_t)(p * 4294967296.0) & ~3u;
        }
        CU(cudaMemcpyToSymbol(c_sub_thresh, th, sizeof th));
        table_ready[c->device & 63] = true;
    }
    cudaStream_t st = c->stream;
    k_synth_windows<<<d.n_windows, 128, (size_t)db->max_hap_len + 32, st>>>(d, seed, first_window);
    if ((rc = launch_check(c, "k_synth_windows"))) return rc;
    k_synth_reads<<<(unsigned)((d.n_slots * 32 + 127) / 128), 128, 0, st>>>(d, seed, first_window);
    return launch_check(c, "k_synth_reads");
}

// Copies the INPUT arrays of a resident batch back into a host batch of the same shape (the batch that was uploaded, or
// one with identical offsets): sequences, qualities, read fields, window coordinates, variant masks and priors.
extern "C" int plb_batch_download(PlbContext* c, PlbDeviceBatch* db, PlbWindowBatch* hb) {
    if (!c || !db || !hb) return set_err(PLB_ERR_ARG, "NULL argument");
    const DevBatch& d = db->d;
    if (hb->n_windows != d.n_windows || hb->n_haps != d.n_haps || hb->n_reads != d.n_reads || hb->n_slots != d.n_slots ||
        hb->seq_format != PLB_SEQ_ASCII || hb->qual_bits != 0)
        return set_err(PLB_ERR_ARG, "plb_batch_download: the host batch must have the shape of the resident one (ASCII)");
    CU(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const int W = d.n_windows;
    auto dn = [&](const void* dst, const void* src, size_t bytes) -> cudaError_t {
        if (!dst || !src || !bytes) return cudaSuccess;
        return cudaMemcpyAsync((void*)dst, src, bytes, cudaMemcpyDeviceToHost, st);
    };
    CU(dn(hb->hap_seq, d.hap_seq, (size_t)db->hap_bytes));
    CU(dn(hb->read_seq, d.read_seq, (size_t)db->read_bytes));
    CU(dn(hb->read_qual, d.read_qual, (size_t)db->read_bytes));
    CU(dn(hb->read_pos, d.read_pos, (size_t)d.n_reads * 4));
    CU(dn(hb->read_end, d.read_end, (size_t)d.n_reads * 4));
    CU(dn(hb->read_mapq, d.read_mapq, (size_t)d.n_reads));
    CU(dn(hb->read_qcfail, d.read_qcfail, (size_t)d.n_reads));
    CU(dn(hb->win_start, d.win_start, (size_t)W * 4));
    CU(dn(hb->win_end, d.win_end, (size_t)W * 4));
    CU(dn(hb->hap_start, d.hap_start, (size_t)W * 4));
    if (db->have_var && hb->max_variants == d.max_variants) {
        CU(dn(hb->win_n_var, d.win_n_var, (size_t)W * 4));
        CU(dn(hb->hap_var_mask, d.hap_var_mask, (size_t)d.n_haps * 8));
        CU(dn(hb->var_prior, d.var_prior, (size_t)W * d.max_variants * 8));
    }
    CU(cudaStreamSynchronize(st));
    return PLB_OK;
}
