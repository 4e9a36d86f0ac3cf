// Microbenchmark: tcgen05.mma issue/execute rate vs N, number of independent accumulators,
// SS (A from smem) vs TS (A from tmem).  One CTA per SM; cycles per MMA printed.
#include <cstdio>
#include <cuda_runtime.h>
#include "../normalizing-flows_b200/csrc/nfb_common.cuh"
void nfb_set_error(const char*, ...) {}
using namespace nfb;

template <int N, int NACC, bool TS, int KSTEP_ADV>
__global__ void __launch_bounds__(128, 1) bench(long long* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    __shared__ uint32_t tptr;
    __shared__ __align__(8) uint64_t barmem;
    const uint32_t bar = smem_u32(&barmem);
    for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tptr;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x < 32) {
        const uint64_t ad = umma_desc_sw128(sbase);
        const uint64_t bd = umma_desc_sw128(sbase + 16384);
        const uint32_t idesc = umma_idesc_bf16(128, N);
        t0 = clock64();
        if (elect_one_sync()) {
            for (int it = 0; it < iters; ++it) {
#pragma unroll
                for (int a = 0; a < NACC; ++a) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint32_t d = tmem + (TS ? 256 : 0) + a * ((TS ? 256 : 512) / NACC);
                        const int k = KSTEP_ADV ? ks : 0;
                        if (TS) umma_bf16_ts(d, tmem + k * 8, bd + 2 * k, idesc, 1u);
                        else umma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, 1u);
                    }
                }
            }
            umma_commit(bar);
        }
        __syncwarp();
        mbar_wait(bar, 0, nullptr, 0);
        t1 = clock64();
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

template <int N, int NACC, bool TS>
void run(const char* tag, int grid) {
    long long* d;
    cudaMalloc(&d, 8 * 256);
    const int iters = 64;
    auto k = bench<N, NACC, TS, 1>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    k<<<grid, 128, 65536>>>(d, iters);
    k<<<grid, 128, 65536>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[256];
    cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
    const double per = (double)mx / (iters * NACC * 4);
    printf("%s N=%3d nacc=%d %s grid=%3d: %7.1f cycles/MMA (ideal %5.1f)  eff %.2f  %s\n", tag, N, NACC,
           TS ? "TS" : "SS", grid, per, N / 2.0, (N / 2.0) / per, e == cudaSuccess ? "" : cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    for (int grid : {1, 148}) {
        run<64, 1, false>("", grid);  run<96, 1, false>("", grid);  run<128, 1, false>("", grid);
        run<192, 1, false>("", grid); run<256, 1, false>("", grid);
        run<64, 2, false>("", grid);  run<96, 2, false>("", grid);  run<128, 2, false>("", grid);
        run<256, 2, false>("", grid); run<64, 4, false>("", grid);  run<128, 4, false>("", grid);
        run<64, 1, true>("", grid);   run<96, 1, true>("", grid);   run<128, 1, true>("", grid);
        run<256, 1, true>("", grid);  run<96, 2, true>("", grid);   run<128, 2, true>("", grid);
        run<64, 4, true>("", grid);
    }
    return 0;
}
