// Semantics test + micro-benchmark for tcgen05 CTA pairs (cta_group::2), the next structural step for the fused
// kernel.  NOT product code; compile: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a
// -o tools/pair_mma_test tools/pair_mma_test.cu ; run on a B200: prints PASS/FAIL per question and cycles/MMA.
//
// Questions it answers (each printed separately so that a wrong assumption is visible, not fatal):
//   Q1  M=256 / cta_group::2: D rows 0..127 land in CTA 0's TMEM lanes, rows 128..255 in CTA 1's, when each CTA
//       supplies ITS 128 rows of A and HALF of B's rows (N/2 x K) at the SAME shared-memory offsets.
//   Q2  which half: CTA r must hold B rows [r N/2, (r+1) N/2).
//   Q3  tcgen05.commit...multicast::cluster with mask 0b11 arrives on the barrier at the same offset in both CTAs.
//   Q4  a remote mbarrier arrive (mapa + mbarrier.arrive.shared::cluster) from CTA 1 is seen by CTA 0's waiter.
//   Q5  throughput: cycles per M=256,N,K=16 MMA with both operands in shared memory (SS) vs cta_group::1 N/2.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../normalizing-flows_b200/csrc/nfb_common.cuh"
void nfb_set_error(const char*, ...) {}
using namespace nfb;

__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void umma2_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t sw128_off(int r, int k) {  // byte offset of element (row r, k) in a SW128 tile
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((k >> 3) ^ (r & 7))) << 4) + (k & 7) * 2;
}

struct Out {
    float d[2][128][256];  // [cta][row][col] accumulator read back
    int commit_seen[2];
    int remote_seen;
    long long cycles2, cycles1;
};

// A[m][k] = ((m * 3 + k) % 7) - 3 ;  B[n][k] = ((n * 3 + 2 k) % 5) - 2   (exact in bf16, sums exact in fp32)
__host__ __device__ inline float a_val(int m, int k) { return (float)(((m * 3 + k) % 7) - 3); }
__host__ __device__ inline float b_val(int n, int k) { return (float)(((n * 3 + 2 * k) % 5) - 2); }

template <int N, bool SWAP_HALVES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) pair_kernel(Out* out, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t rank = cluster_ctarank();
    __shared__ uint32_t tptr;
    __shared__ __align__(8) uint64_t bars[4];  // 0: commit (multicast), 1: remote-arrive test, 2: timing commit
    // A tile: this CTA's 128 rows (global rows rank*128 + r) x K=64 ; B half: N/2 rows x 64
    uint8_t* At = smem;
    uint8_t* Bt = smem + 16384;
    for (int i = threadIdx.x; i < 128 * 64; i += 128) {
        const int r = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(At + sw128_off(r, k)) = __float2bfloat16(a_val((int)rank * 128 + r, k));
    }
    const int half = SWAP_HALVES ? 1 - (int)rank : (int)rank;
    for (int i = threadIdx.x; i < (N / 2) * 64; i += 128) {
        const int n = i >> 6, k = i & 63;
        *reinterpret_cast<__nv_bfloat16*>(Bt + sw128_off(n, k)) = __float2bfloat16(b_val(half * (N / 2) + n, k));
    }
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars[0]), 1);
        mbar_init(smem_u32(&bars[1]), 1);
        mbar_init(smem_u32(&bars[2]), 1);
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    __syncthreads();
    cluster_sync();  // barriers of both CTAs initialised before anyone signals them
    if (threadIdx.x < 32) {  // one warp of EACH CTA takes part in the pair allocation
        tmem_alloc2(smem_u32(&tptr), 512);
        tmem_relinquish2();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tptr;
    cluster_sync();

    // ---- Q4: remote arrive ----
    if (rank == 1 && threadIdx.x == 0) mbar_arrive_remote(mapa(smem_u32(&bars[1]), 0));
    if (rank == 0 && threadIdx.x == 0) {
        mbar_wait(smem_u32(&bars[1]), 0, nullptr, 0);
        out->remote_seen = 1;
    }

    // ---- Q1-Q3: one K=64 product ----
    const uint32_t idesc2 = umma_idesc_bf16(256, N);
    if (rank == 0 && threadIdx.x < 32) {
        if (elect_one_sync()) {
            const uint64_t ad = umma_desc_sw128(sbase), bd = umma_desc_sw128(sbase + 16384);
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) umma2_bf16(tmem, ad + 2 * ks, bd + 2 * ks, idesc2, ks ? 1u : 0u);
            umma2_commit_mc(smem_u32(&bars[0]), 0b11);
        }
        __syncwarp();
    }
    if (threadIdx.x == 0) {
        mbar_wait(smem_u32(&bars[0]), 0, nullptr, 0);  // both CTAs wait on their OWN barrier
        out->commit_seen[rank] = 1;
    }
    __syncthreads();
    tc_fence_after();
    {
        const int w = threadIdx.x >> 5, lane = threadIdx.x & 31, r = w * 32 + lane;
        const uint32_t tl = tmem + ((uint32_t)(w * 32) << 16);
        for (int n0 = 0; n0 < N; n0 += 16) {
            uint32_t v[16];
            NFB_TMEM_LD16(tl + n0, v);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) out->d[rank][r][n0 + j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();

    // ---- Q5: throughput, cta_group::2 (M=256, N) ----
    if (rank == 0 && threadIdx.x < 32) {
        const long long t0 = clock64();
        if (elect_one_sync()) {
            const uint64_t ad = umma_desc_sw128(sbase), bd = umma_desc_sw128(sbase + 16384);
            for (int it = 0; it < iters; ++it)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma2_bf16(tmem, ad + 2 * ks, bd + 2 * ks, idesc2, 1u);
            umma2_commit_mc(smem_u32(&bars[2]), 0b11);
        }
        __syncwarp();
        mbar_wait(smem_u32(&bars[2]), 0, nullptr, 0);
        if (threadIdx.x == 0) out->cycles2 = clock64() - t0;
    } else if (threadIdx.x == 0) {
        mbar_wait(smem_u32(&bars[2]), 0, nullptr, 0);
    }
    __syncthreads();
    cluster_sync();
    if (threadIdx.x < 32) tmem_dealloc2(tmem, 512);
}

template <int N, bool SWAP>
bool run(const char* tag) {
    Out* d;
    cudaMalloc(&d, sizeof(Out));
    cudaMemset(d, 0, sizeof(Out));
    auto k = pair_kernel<N, SWAP>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
    const int iters = 256;
    k<<<2, 128, 65536>>>(d, iters);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s N=%d swap=%d: CUDA error %s\n", tag, N, (int)SWAP, cudaGetErrorString(e));
        cudaFree(d);
        return false;
    }
    std::vector<char> hb(sizeof(Out));
    cudaMemcpy(hb.data(), d, sizeof(Out), cudaMemcpyDeviceToHost);
    const Out& h = *reinterpret_cast<const Out*>(hb.data());
    long long bad = 0;
    double maxerr = 0;
    for (int c = 0; c < 2; ++c)
        for (int r = 0; r < 128; ++r)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int kk = 0; kk < 64; ++kk) ref += (double)a_val(c * 128 + r, kk) * b_val(n, kk);
                const double err = fabs(ref - h.d[c][r][n]);
                maxerr = err > maxerr ? err : maxerr;
                bad += err > 1e-3;
            }
    printf("%s N=%3d B-half-of-CTA-r=%s: D %s (bad %lld, max err %.3g) | commit seen cta0=%d cta1=%d | remote arrive %d | "
           "%.1f cycles/MMA (M=256; cta_group::1 at N/2 per SM would be %.0f)\n",
           tag, N, SWAP ? "rows of the OTHER rank" : "rows [r N/2, (r+1) N/2)", bad == 0 ? "PASS" : "FAIL", bad, maxerr,
           h.commit_seen[0], h.commit_seen[1], h.remote_seen, (double)h.cycles2 / (iters * 4), N / 2.0);
    cudaFree(d);
    return bad == 0;
}

int main() {
    int dev = 0, major = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (major != 10) { printf("needs sm_100\n"); return 1; }
    bool ok = false;
    ok |= run<256, false>("pair");
    ok |= run<256, true>("pair");
    run<128, false>("pair");
    run<240, false>("pair");
    run<80, false>("pair");
    printf(ok ? "at least one operand placement reproduces A B^T\n" : "NO placement matched: re-read the PTX ISA\n");
    return 0;
}
