// scripts/ubench/tma_probe.cu -- what does cp.async.bulk.tensor.3d do with a uint8 box that starts at negative
// coordinates / hangs over the right and bottom edge?  (k_blur stages its halo tiles that way.)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <vector>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int BW, int BH>
__global__ void probe(const __grid_constant__ CUtensorMap map, int x, int y, int f, uint8_t *out, int *status)
{
    __shared__ __align__(128) uint8_t box[BH][BW];
    __shared__ __align__(8) uint64_t mbar;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(1) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(&mbar)), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                     :: "r"(smem_u32(&box[0][0])), "l"(reinterpret_cast<unsigned long long>(&map)), "r"(smem_u32(&mbar)),
                        "r"(x), "r"(y), "r"(f) : "memory");
    }
    int ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin) {
        unsigned o;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(o) : "r"(smem_u32(&mbar)), "r"(0) : "memory");
        ok = (int)o;
    }
    if (threadIdx.x == 0) *status = ok;
    __syncthreads();
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = (&box[0][0])[i];
}

int main()
{
    const int pitch = 1280, W = 1242, H = 375, F = 2, BW = 144, BH = 70;
    std::vector<uint8_t> img((size_t)pitch * H * F);
    for (int f = 0; f < F; ++f) for (int y = 0; y < H; ++y) for (int x = 0; x < pitch; ++x) img[((size_t)f * H + y) * pitch + x] = (uint8_t)(1 + (x * 7 + y * 13 + f * 101) % 251);
    uint8_t *d_img, *d_out; int *d_status;
    cudaMalloc(&d_img, img.size()); cudaMalloc(&d_out, BW * BH); cudaMalloc(&d_status, 4);
    cudaMemcpy(d_img, img.data(), img.size(), cudaMemcpyHostToDevice);
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    encode_fn fn = (encode_fn)p;
    for (int variant = 0; variant < 2; ++variant) {
        CUtensorMap map;
        // variant 0: dim0 = pitch (what the library does); variant 1: dim0 = image width
        const cuuint64_t gdim[3] = {(cuuint64_t)(variant ? W : pitch), (cuuint64_t)H, (cuuint64_t)F};
        const cuuint64_t gstr[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * H};
        const cuuint32_t box[3] = {BW, BH, 1}, estr[3] = {1, 1, 1};
        CUresult r = fn(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_img, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("variant %d encode -> %d\n", variant, (int)r);
        const int coords[][3] = {{0, 0, 0}, {16, 5, 1}, {112, 61, 0}, {-16, 0, 0}, {0, -3, 0}, {-16, -3, 0}, {1136, 317, 1}, {1264, 372, 1}, {-32, -70, 0}, {1280, 0, 0}};
        for (auto &c : coords) {
            cudaMemset(d_out, 0xEE, BW * BH); cudaMemset(d_status, 0xFF, 4);
            probe<BW, BH><<<1, 128>>>(map, c[0], c[1], c[2], d_out, d_status);
            cudaError_t e = cudaDeviceSynchronize();
            int st = -1; std::vector<uint8_t> out(BW * BH);
            if (e == cudaSuccess) { cudaMemcpy(&st, d_status, 4, cudaMemcpyDeviceToHost); cudaMemcpy(out.data(), d_out, BW * BH, cudaMemcpyDeviceToHost); }
            long bad = 0;
            const int wlim = variant ? W : pitch;
            for (int r2 = 0; r2 < BH; ++r2) for (int cc = 0; cc < BW; ++cc) {
                const int gx = c[0] + cc, gy = c[1] + r2;
                const uint8_t want = (gx < 0 || gy < 0 || gx >= wlim || gy >= H) ? 0 : img[((size_t)c[2] * H + gy) * pitch + gx];
                bad += out[r2 * BW + cc] != want;
            }
            printf("  box at (%d,%d,%d): err=%s completed=%d mismatches=%ld\n", c[0], c[1], c[2], cudaGetErrorString(e), st, bad);
            if (e != cudaSuccess) { printf("  context is dead, stopping\n"); return 1; }
        }
    }
    return 0;
}
