// Micro-benchmark: which issue pipes do VIMNMX3.S16x2 / HMNMX2 / VIADD.16x2 / IMAD use on sm_100a?
// Each kernel runs independent dependency chains; mixed kernels tell whether two instruction kinds overlap.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define N 4096
template <int MODE> __global__ void k(unsigned *out, unsigned seed)
{
    unsigned a0 = seed + threadIdx.x, a1 = a0 * 3, a2 = a0 * 5, a3 = a0 * 7, b0 = a0 ^ 0x55, b1 = a1 ^ 0x33, b2 = a2 ^ 0x11, b3 = a3 ^ 0x77;
    __half2 h0 = *(__half2 *)&b0, h1 = *(__half2 *)&b1, h2 = *(__half2 *)&b2, h3 = *(__half2 *)&b3;
    const __half2 hc = __floats2half2_rn(0.5f, 0.25f);
    float f0 = __uint_as_float(a0 & 0x3fffffff), f1 = __uint_as_float(a1 & 0x3fffffff), f2 = __uint_as_float(a2 & 0x3fffffff), f3 = __uint_as_float(a3 & 0x3fffffff);
    const float fs = __uint_as_float(seed);
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (MODE == 0 || MODE == 2 || MODE == 4) { a0 = __vimin3_s16x2(a0, a1, seed); a1 = __vimax3_s16x2(a1, a2, seed); a2 = __vimin3_s16x2(a2, a3, seed); a3 = __vimax3_s16x2(a3, a0, seed); }
        if (MODE == 1 || MODE == 2) { h0 = __hmin2(h0, h1); h1 = __hmax2(h1, h2); h2 = __hmin2(h2, h3); h3 = __hmax2(h3, hc); }
        if (MODE == 3 || MODE == 4) { h0 = __hfma2(h0, hc, h1); h1 = __hfma2(h1, hc, h2); h2 = __hfma2(h2, hc, h3); h3 = __hfma2(h3, hc, h0); }
        if (MODE == 5) { a0 = __vadd2(a0, a1); a1 = __vadd2(a1, a2); a2 = __vadd2(a2, a3); a3 = __vadd2(a3, seed); }
        if (MODE == 6) { a0 = min(a0, a1); a1 = max(a1, a2); a2 = min(a2, a3); a3 = max(a3, seed); }
        if (MODE == 7 || MODE == 8) { f0 = fminf(f0, f1); f1 = fmaxf(f1, f2); f2 = fminf(f2, f3); f3 = fmaxf(f3, fs); }
        if (MODE == 8 || MODE == 10) { a0 = __vimin3_s16x2(a0, a1, seed); a1 = __vimax3_s16x2(a1, a2, seed); a2 = __vimin3_s16x2(a2, a3, seed); a3 = __vimax3_s16x2(a3, a0, seed); }
        if (MODE == 9 || MODE == 10) { b0 = b0 * seed + b1; b1 = b1 * seed + b2; b2 = b2 * seed + b3; b3 = b3 * seed + b0; }
        if (MODE == 11) { a0 = __vimax3_s32(a0, a1, seed); a1 = __vimin3_s32(a1, a2, seed); a2 = __vimax3_s32(a2, a3, seed); a3 = __vimin3_s32(a3, a0, seed); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ *(unsigned *)&h0 ^ *(unsigned *)&h1 ^ *(unsigned *)&h2 ^ *(unsigned *)&h3 ^ __float_as_uint(f0 + f1 + f2 + f3) ^ b0 ^ b1 ^ b2 ^ b3;
}
template <int MODE> float run(unsigned *d)
{
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 8, 256>>>(d, 1234); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, 1234); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main()
{
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const char *names[] = {"VIMNMX3.S16x2 x4", "HMNMX2 x4", "VIMNMX3 x4 + HMNMX2 x4", "HFMA2 x4", "VIMNMX3 x4 + HFMA2 x4", "VIADD.16x2 x4", "IMNMX x4",
                           "FMNMX x4", "VIMNMX3 x4 + FMNMX x4", "IMAD x4", "VIMNMX3 x4 + IMAD x4", "VIMNMX3.S32 x4"};
    float t[12] = {run<0>(d), run<1>(d), run<2>(d), run<3>(d), run<4>(d), run<5>(d), run<6>(d), run<7>(d), run<8>(d), run<9>(d), run<10>(d), run<11>(d)};
    const double warp_instr = 148.0 * 8 * 8 * N * 4;      // per 4-op group
    for (int i = 0; i < 12; ++i) printf("%-28s %.3f ms  -> %.2f warp-instr/clk/SM (at 1.965 GHz, counting 4 ops/iter%s)\n", names[i], t[i],
                                      warp_instr * ((i == 2 || i == 4 || i == 8 || i == 10) ? 2 : 1) / (t[i] * 1e-3) / 1.965e9 / 148, (i == 2 || i == 4 || i == 8 || i == 10) ? " x2" : "");
    return 0;
}
