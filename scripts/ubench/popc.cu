// Micro-benchmark: peak POPC.32 rate on this GPU (the matcher's roofline is the integer pipe, not HBM).
// Independent chains of POPC feeding an add; also the matcher's real inner mix (XOR + carry-save LOP3 + 5 POPC + adds per pair).
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
template <int MODE> __global__ void k(unsigned *out, unsigned seed)
{
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, a4 = a0 * 11, a5 = a0 * 13, a6 = a0 * 17, a7 = a0 * 19;
    unsigned s = 0;
#pragma unroll 8
    for (int i = 0; i < N; ++i) {
        if (MODE == 0) {                                                    // 8 independent POPC per iteration
            a0 = __popc(a0 ^ seed) + i; a1 = __popc(a1 ^ seed) + i; a2 = __popc(a2 ^ seed) + i; a3 = __popc(a3 ^ seed) + i;
            a4 = __popc(a4 ^ seed) + i; a5 = __popc(a5 ^ seed) + i; a6 = __popc(a6 ^ seed) + i; a7 = __popc(a7 ^ seed) + i;
        } else {                                                            // one 256-bit pair the way k_match_partial does it
            const unsigned b = seed * (i + 1);
            const unsigned x0 = a0 ^ b, x1 = a1 ^ (b + 1), x2 = a2 ^ (b + 2), x3 = a3 ^ (b + 3), x4 = a4 ^ (b + 4), x5 = a5 ^ (b + 5), x6 = a6 ^ (b + 6), x7 = a7 ^ (b + 7);
            const unsigned s0 = x0 ^ x1 ^ x2, c0 = (x0 & x1) | (x2 & (x0 | x1));
            const unsigned s1 = x3 ^ x4 ^ x5, c1 = (x3 & x4) | (x5 & (x3 | x4));
            const unsigned s2 = s0 ^ s1 ^ x6, c2 = (s0 & s1) | (x6 & (s0 | s1));
            s += __popc(s2) + __popc(x7) + 2 * (__popc(c0) + __popc(c1) + __popc(c2));
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ s;
}
template <int MODE> float run(unsigned *d)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1234); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, 1234); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
    }
    return best;
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const float t0 = run<0>(d), t1 = run<1>(d);
    const double thr = 148.0 * 8 * 256 * (double)N;                         // thread-iterations
    printf("{\"popc32_per_s\": %.4e, \"popc32_per_clk_per_sm\": %.2f, \"matcher_mix_pairs_per_s\": %.4e, \"matcher_mix_popc32_equiv_per_s\": %.4e}\n",
           thr * 8 / (t0 * 1e-3), thr * 8 / (t0 * 1e-3) / 1.965e9 / 148, thr / (t1 * 1e-3), thr * 8 / (t1 * 1e-3));
    return 0;
}
