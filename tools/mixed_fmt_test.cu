// Does tcgen05.mma.kind::f16 accept A=bf16 with B=fp16 (and vice versa) in one instruction?
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../normalizing-flows_b200/csrc/nfb_common.cuh"
void nfb_set_error(const char*, ...) {}
using namespace nfb;
__host__ __device__ constexpr uint32_t idesc_fmt(uint32_t afmt, uint32_t bfmt, uint32_t m, uint32_t n) {
    return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);  // fmt: 0 f16, 1 bf16
}
__global__ void __launch_bounds__(128, 1) test(float* out, int afmt, int bfmt) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tptr;
    __shared__ __align__(8) uint64_t barmem;
    const uint32_t sbase = smem_u32(smem), bar = smem_u32(&barmem);
    // A: 128 rows, row r filled with value (1 + r/256) in format afmt; B: 64 rows, row n filled with (0.5 + n/128)
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        const int r = i / 64;
        const float v = 1.0f + r / 256.0f;
        uint16_t bits = afmt ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
        reinterpret_cast<uint16_t*>(smem)[(r / 8) * 512 + (r % 8) * 64 + (i % 64)] = bits;
    }
    for (int i = threadIdx.x; i < 64 * 64; i += 128) {
        const int n = i / 64;
        const float v = 0.5f + n / 128.0f;
        uint16_t bits = bfmt ? __bfloat16_as_ushort(__float2bfloat16_rn(v)) : __half_as_ushort(__float2half_rn(v));
        reinterpret_cast<uint16_t*>(smem + 16384)[(n / 8) * 512 + (n % 8) * 64 + (i % 64)] = bits;
    }
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 64); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tptr;
    if (threadIdx.x < 32) {
        if (elect_one_sync()) {
            umma_bf16(tmem, umma_desc_sw128(sbase), umma_desc_sw128(sbase + 16384), idesc_fmt(afmt, bfmt, 128, 64), 0u);
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0, nullptr, 0);
    tc_fence_after();
    uint32_t v[8];
    const int q = threadIdx.x >> 5;
    NFB_TMEM_LD8(tmem + ((uint32_t)(q * 32) << 16), v);
    tc_wait_ld();
    for (int j = 0; j < 8; ++j) out[threadIdx.x * 8 + j] = __uint_as_float(v[j]);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 64);
}
int main() {
    float* d; cudaMalloc(&d, 128 * 8 * 4);
    float h[1024];
    for (int afmt = 0; afmt < 2; ++afmt) for (int bfmt = 0; bfmt < 2; ++bfmt) {
        cudaFuncSetAttribute(test, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
        test<<<1, 128, 32768>>>(d, afmt, bfmt);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int r = 0; r < 128; ++r) for (int n = 0; n < 8; ++n) {
            const double ref = 16.0 * (1.0 + r / 256.0) * (0.5 + n / 128.0);  // K=16 identical products (values exact in both formats)
            maxerr = fmax(maxerr, fabs(h[r * 8 + n] - ref));
        }
        printf("A=%s B=%s: %s  D[5][3]=%f (ref %f)  max err %.3e\n", afmt ? "bf16" : "f16", bfmt ? "bf16" : "f16",
               e == cudaSuccess ? "ok" : cudaGetErrorString(e), h[5 * 8 + 3], 16.0 * (1 + 5 / 256.0) * (0.5 + 3 / 128.0), maxerr);
        if (e != cudaSuccess) break;
    }
    return 0;
}
