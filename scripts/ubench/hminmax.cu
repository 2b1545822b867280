// Micro-benchmark: packed 16x2 min / max of small non-negative integers on the FMA pipe.
// For 0 <= n < 2048 the int16 bit pattern n IS the fp16 value n * 2^-24 (denormal / first normal binade), and sums / differences of
// such values are exact in fp16, so   d = relu(a - b)  (HFMA2.RELU),  max = b + d,  min = a - d  (HADD2)
// gives both results of a pair in three FMA-pipe instructions with the integer bit patterns intact -- work the integer ALU pipe
// (VIMNMX, 64 lanes / clk / SM) does not have to do.  Checks exactness over all byte pairs, then times the mixes.
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned relu_sub(unsigned a, unsigned b)      // relu(a - b) per half
{
    unsigned d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(0xBC00BC00u), "r"(a));   // b * -1.0 + a
    return d;
}
__device__ __forceinline__ unsigned hadd(unsigned a, unsigned b) { unsigned d; asm("add.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
__device__ __forceinline__ unsigned hsub(unsigned a, unsigned b) { unsigned d; asm("sub.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }

__global__ void k_check(int *bad)
{
    const int a = blockIdx.x, b = threadIdx.x;                              // all byte pairs, both halves at once with different values
    const unsigned pa = a | (255 - a) << 16, pb = b | ((b * 7) & 255) << 16;
    const unsigned d = relu_sub(pa, pb), mx = hadd(pb, d), mn = hsub(pa, d);
    const unsigned emx = max(a, b) | max(255 - a, (b * 7) & 255) << 16, emn = min(a, b) | min(255 - a, (b * 7) & 255) << 16;
    if (mx != emx || mn != emn) atomicAdd(bad, 1);
}

#define N 4096
template <int MODE> __global__ void k(unsigned *out, unsigned seed)
{
    unsigned a0 = (seed + threadIdx.x) & 0x00ff00ff, a1 = (a0 * 3) & 0x00ff00ff, a2 = (a0 * 5) & 0x00ff00ff, a3 = (a0 * 7) & 0x00ff00ff;
    unsigned b0 = a0 ^ 0x55, b1 = a1 ^ 0x33, b2 = a2 ^ 0x11, b3 = a3 ^ 0x77;
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (MODE == 0 || MODE == 2) { a0 = __vimin3_s16x2(a0, a1, seed); a1 = __vimax3_s16x2(a1, a2, seed); a2 = __vimin3_s16x2(a2, a3, seed); a3 = __vimax3_s16x2(a3, a0, seed); }
        if (MODE == 1 || MODE == 2) {                                       // two (min, max) pairs = six FMA-pipe instructions
            const unsigned d0 = relu_sub(b0, b1), d1 = relu_sub(b2, b3);
            const unsigned mx0 = hadd(b1, d0), mn0 = hsub(b0, d0), mx1 = hadd(b3, d1), mn1 = hsub(b2, d1);
            b0 = mx0; b1 = mn1; b2 = mx1; b3 = mn0;
        }
        if (MODE == 3) { a0 = __vmins2(a0, a1); a1 = __vmaxs2(a1, a2); a2 = __vmins2(a2, a3); a3 = __vmaxs2(a3, seed); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
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
    int *bad; cudaMallocManaged(&bad, 4); *bad = 0;
    k_check<<<256, 256>>>(bad); cudaDeviceSynchronize();
    printf("fp16-denormal min/max of all byte pairs: %d mismatches\n", *bad);
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    const float t0 = run<0>(d), t1 = run<1>(d), t2 = run<2>(d), t3 = run<3>(d);
    const double w = 148.0 * 8 * 8 * N;                                     // warps x iterations
    auto rate = [&](double instr_per_iter, float ms) { return w * instr_per_iter / (ms * 1e-3) / 1.965e9 / 148; };
    printf("VIMNMX3.S16x2 x4           %.3f ms  %.2f warp-instr/clk/SM\n", t0, rate(4, t0));
    printf("HFMA2.RELU x2 + HADD2 x4   %.3f ms  %.2f warp-instr/clk/SM  (= %.2f packed min+max pairs/clk/SM)\n", t1, rate(6, t1), rate(2, t1));
    printf("both in one loop           %.3f ms  %.2f warp-instr/clk/SM  (sum of the two alone: %.3f ms)\n", t2, rate(10, t2), t0 + t1);
    printf("VIMNMX.S16x2 (2-input) x4  %.3f ms  %.2f warp-instr/clk/SM\n", t3, rate(4, t3));
    return 0;
}
