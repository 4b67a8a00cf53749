// microbench.cu — issue-rate micro-benchmark for the packed-halfword instructions the band DP uses
// (VIADD.16x2, VIMNMX.S16x2, VIADDMNMX.S16x2, PRMT) plus IADD3 / IMAD / LOP3 for comparison, and the
// exact 8-op cell-update mix of plb_dp.cuh.  Prints warp-instructions per cycle per SM.
// This is the "second yardstick" next to the HBM roofline: the path is integer-issue
// bound, so kernel quality is (achieved issue rate) / (measured peak issue rate of this mix).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

typedef uint32_t u32;
constexpr int ITERS = 4096;
constexpr int CH = 8;  // independent chains per thread

__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 s) {
    u32 d;
    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(s));
    return d;
}

template <int OP>
__global__ void __launch_bounds__(256) k(u32* out, long long* cyc, u32 seed) {
    u32 x[CH], y[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) {
        x[i] = seed * (threadIdx.x + 1) + i * 0x9E37u;
        y[i] = seed ^ (i * 0x85EBu + threadIdx.x);
    }
    const u32 c1 = seed | 0x00030003u, c2 = seed & 0x000F000Fu;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (OP == 0) { x[i] = __vadd2(x[i], y[i]); y[i] = __vadd2(y[i], c1); }
            if (OP == 1) { x[i] = __vmins2(x[i], y[i]); y[i] = __vadd2(y[i], c1); }        // 1 min + 1 add
            if (OP == 2) { x[i] = __viaddmin_s16x2(x[i], c1, y[i]); y[i] = __vadd2(y[i], c2); }  // 1 addmin + 1 add
            if (OP == 3) { x[i] = prmt(x[i], y[i], c1); y[i] = __vadd2(y[i], c2); }        // 1 prmt + 1 add
            if (OP == 4) { x[i] = x[i] + y[i] + c1; y[i] = y[i] + x[i] + c2; }             // IADD3 x2
            if (OP == 5) { x[i] = x[i] * c1 + y[i]; y[i] = y[i] * c2 + x[i]; }             // IMAD x2
            if (OP == 6) { x[i] = (x[i] & y[i]) ^ c1; y[i] = (y[i] | x[i]) ^ c2; }         // LOP3 x2
            if (OP == 7) { x[i] = __vimin3_s16x2(x[i], y[i], c1); y[i] = __vadd2(y[i], c2); }  // 1 min3 + 1 add
            if (OP == 8) {  // the cell update of plb_dp.cuh: min, prmt, add, add, addmin, add, addmin, min
                u32 B = __vmins2(x[i], y[i]);
                u32 sub = prmt(c1, c2, y[i]);
                u32 Mn = __vadd2(B, sub);
                u32 In = __viaddmin_s16x2(y[i], c1, __vadd2(x[i], c2));
                u32 Dn = __viaddmin_s16x2(x[i], c2, __vadd2(y[i], c1));
                x[i] = __vmins2(Mn, In);
                y[i] = Dn;
            }
            if (OP == 10) {  // 6-op cell update (frame-shifted): 2 min, 2 addmin, prmt, 1 add
                u32 B = __vmins2(x[i], y[i]);
                u32 Mn = __vadd2(B, prmt(c1, c2, y[i]));
                u32 In = __viaddmin_s16x2(y[i], c1, x[i]);
                u32 Dn = __viaddmin_s16x2(x[i], c2, y[i]);
                x[i] = __vmins2(Mn, In);
                y[i] = Dn;
            }
            if (OP == 11) {  // 7-op: prmt, min, 2 min3 (4 ALU-pipe) + 3 adds
                u32 B = __vimin3_s16x2(x[i], y[i], c2);
                u32 Mn = __vadd2(B, prmt(c1, c2, y[i]));
                u32 In = __viaddmin_s16x2(y[i], c1, x[i]);
                u32 Dn = __vimin3_s16x2(__vadd2(x[i], c2), __vadd2(y[i], c2), Mn);
                x[i] = In ^ Mn;   // keep both live with one cheap op (LOP3 counted in the 8)
                y[i] = Dn;
            }
            if (OP == 12) { x[i] = __vimin3_s16x2(x[i], y[i], c1); y[i] = __vimin3_s16x2(y[i], x[i], c2); }  // min3 x2
            if (OP == 13) { x[i] = prmt(x[i], y[i], c1); y[i] = prmt(y[i], x[i], c2); }                    // prmt x2
            if (OP == 14) { x[i] = __viaddmin_s16x2(x[i], c1, y[i]); y[i] = __viaddmin_s16x2(y[i], c2, x[i]); }  // addmin x2
            if (OP == 9) {  // same mix with the three plain adds as 32-bit IMAD (fma pipe); lanes never carry
                u32 B = __vmins2(x[i], y[i]);
                u32 sub = prmt(c1, c2, y[i]);
                u32 Mn = B * 1u + sub;
                u32 t1, t2;
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(t1) : "r"(x[i]), "r"(c2));
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(t2) : "r"(y[i]), "r"(c1));
                asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(Mn) : "r"(B), "r"(sub));
                u32 In = __viaddmin_s16x2(y[i], c1, t1);
                u32 Dn = __viaddmin_s16x2(x[i], c2, t2);
                x[i] = __vmins2(Mn, In);
                y[i] = Dn;
            }
        }
    }
    long long t1 = clock64();
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) acc ^= x[i] ^ y[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_chain, int n_sm, int blocks_per_sm) {
    int grid = n_sm * blocks_per_sm;
    u32* out;
    long long* cyc;
    cudaMalloc(&out, (size_t)grid * 256 * 4);
    cudaMalloc(&cyc, (size_t)grid * 8);
    k<OP><<<grid, 256>>>(out, cyc, 12345u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(out, cyc, 6789u);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    std::vector<long long> h(grid);
    cudaMemcpy(h.data(), cyc, (size_t)grid * 8, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (auto c : h) avg += (double)c;
    avg /= grid;
    double warp_instr_per_sm = (double)blocks_per_sm * 8 /*warps*/ * ITERS * CH * ops_per_chain;
    double ipc = warp_instr_per_sm / avg;
    double ghz = avg / (ms * 1e6);
    double total_wi = (double)grid * 8 * ITERS * CH * ops_per_chain;
    printf("%-44s blocks/SM=%d  %.3f Gwarp-instr/s/SM = %.2f warp-instr/clk/SM @1.965GHz  (kernel %.3f ms; in-kernel %.2f/clk, ~%.2f GHz)\n",
           name, blocks_per_sm, total_wi / (ms * 1e-3) / n_sm / 1e9, total_wi / (ms * 1e-3) / n_sm / 1.965e9, ms, ipc, ghz);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs\n", p.name, p.multiProcessorCount);
    int n = p.multiProcessorCount;
    for (int b : {4, 8}) {
        run<0>("VIADD.16x2 x2", 2, n, b);
        run<1>("VIMNMX.S16x2 + VIADD.16x2", 2, n, b);
        run<2>("VIADDMNMX.S16x2 + VIADD.16x2", 2, n, b);
        run<3>("PRMT + VIADD.16x2", 2, n, b);
        run<4>("IADD3 x2", 2, n, b);
        run<5>("IMAD x2", 2, n, b);
        run<6>("LOP3 x2", 2, n, b);
        run<7>("VIMNMX3.S16x2 + VIADD.16x2", 2, n, b);
        run<8>("DP cell mix (2min,1prmt,3add,2addmin)", 8, n, b);
        run<9>("DP cell mix, adds as IMAD", 8, n, b);
        run<10>("6-op cell (2min,2addmin,prmt,1add)", 6, n, b);
        run<11>("7-op cell (prmt,addmin,2min3,3add,+1lop)", 8, n, b);
        run<12>("VIMNMX3.S16x2 x2", 2, n, b);
        run<13>("PRMT x2", 2, n, b);
        run<14>("VIADDMNMX.S16x2 x2", 2, n, b);
    }
    return 0;
}
