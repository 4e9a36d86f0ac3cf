// nfb_common.cuh -- shared device helpers for the normflows-b200 kernels (sm_100a only).
// PTX wrappers for mbarrier / bulk-copy (TMA) / tcgen05 + error plumbing for the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>

#define NFB_OK 0
#define NFB_ERR_CUDA 1
#define NFB_ERR_ARG 2
#define NFB_ERR_UNSUPPORTED 3
#define NFB_ERR_STATE 4

void nfb_set_error(const char* fmt, ...);

#define NFB_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess) {                                                        \
            nfb_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call,                  \
                          cudaGetErrorString(e__));                                      \
            return NFB_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

#define NFB_CHECK(cond, code, ...)                                                       \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            nfb_set_error(__VA_ARGS__);                                                  \
            return (code);                                                               \
        }                                                                                \
    } while (0)

#define NFB_LAUNCH_CHECK() NFB_CUDA(cudaGetLastError())

namespace nfb {

constexpr float kMinBinWidth = 1e-3f;    // utils/splines.py:6
constexpr float kMinBinHeight = 1e-3f;   // utils/splines.py:7
constexpr float kMinDerivative = 1e-3f;  // utils/splines.py:8
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// One lane of a fully-converged warp.  Issuing tcgen05.mma / cp.async.bulk under this predicate
// (instead of `lane == 0`) lets ptxas keep the operands in uniform registers; with a plain lane test
// it wraps every UTCHMMA in an ELECT/BRA.U.ANY waterfall loop (~17 dependent instructions per MMA).
__device__ __forceinline__ bool elect_one_sync() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 %%rx;\n\t.reg .pred %%px;\n\t"
        "elect.sync %%rx|%%px, %1;\n\t"
        "@%%px mov.s32 %0, 1;\n\t}"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}

// ----------------------------------------------------------------------------------------
// mbarrier
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug must never hang the GPU box.  On timeout the kernel records
// the barrier id in *err and traps; the host sees a launch failure instead of a dead device.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) {  // ~2 s
            if (err) atomicExch(err, tag);
            __threadfence_system();
            asm volatile("trap;");
        }
    }
}

// Warp-level wait: ONE lane polls, the others park at the warp barrier (bar.warp.sync orders lane 0's acquire before
// the other lanes' subsequent accesses).  Used where a whole warp waits for one event.
__device__ __forceinline__ void mbar_wait_warp(uint32_t bar, uint32_t parity, int* err, int tag) {
    if ((threadIdx.x & 31) == 0 && !mbar_try_wait(bar, parity)) {
        long long t0 = clock64();
        while (!mbar_try_wait(bar, parity)) {
            if (clock64() - t0 > 4000000000LL) {
                if (err) atomicExch(err, tag);
                __threadfence_system();
                asm volatile("trap;");
            }
        }
    }
    __syncwarp();
}

// ----------------------------------------------------------------------------------------
// TMA: 1-D bulk copy global -> shared with mbarrier completion (SASS: UBLKCP)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(bar)
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ----------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
                 "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_wait_ld() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]; kind::f16 covers bf16 inputs with fp32 accumulate.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Same, with the A operand read from tensor memory (128 lanes x 8 columns of packed bf16 pairs per
// K=16 step) -- no shared-memory traffic for A.
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc,
                                             uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_wait_st() {
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
#define NFB_TMEM_ST16(addr, v)                                                           \
    asm volatile(                                                                        \
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "                                  \
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"                      \
        ::"r"(addr), "r"((v)[0]), "r"((v)[1]), "r"((v)[2]), "r"((v)[3]), "r"((v)[4]),    \
          "r"((v)[5]), "r"((v)[6]), "r"((v)[7]), "r"((v)[8]), "r"((v)[9]), "r"((v)[10]), \
          "r"((v)[11]), "r"((v)[12]), "r"((v)[13]), "r"((v)[14]), "r"((v)[15])           \
        : "memory")
// mbarrier arrives when all tcgen05.mma issued so far by this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
                 : "memory");
}

// ---- CTA pairs (cluster of 2, tcgen05 cta_group::2): one M=256 MMA spans both SMs; each CTA stages HALF of B ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ uint32_t cluster_nctaid_x() {  // number of clusters along x
    uint32_t r;
    asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address of THIS CTA -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish2() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma2_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair when the MMAs issued so far are done
__device__ __forceinline__ void umma2_commit_mc(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3)
                 : "memory");
}

// K-major, SWIZZLE_128B canonical tile: rows of 128 B (64 bf16), 8-row groups 1024 B apart.
// desc: [0,14) addr>>4 | [16,30) LBO>>4 (=1, ignored for swizzled K-major) | [32,46) SBO>>4 (=64)
//       | [46,48) version=1 (sm_100) | [61,64) layout=2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
    return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | (1ull << 16) | (64ull << 32) | (1ull << 46) |
           (2ull << 61);
}
// fp16 operands (a_format = b_format = 0): the split-precision GEMMs of the fused spline kernel (11-bit mantissas:
// hi + lo carries 22-23 bits of an fp32 value, against 16-17 for a bf16 pair)
__host__ __device__ constexpr uint32_t umma_idesc_f16(uint32_t m, uint32_t n) {
    return (1u << 4) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
// instruction descriptor: c=f32 (bit4), a=bf16 (bit7), b=bf16 (bit10), K-major A and B,
// N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t m, uint32_t n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

#define NFB_TMEM_LD8(addr, v)                                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" \
                 : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]),   \
                   "=r"((v)[5]), "=r"((v)[6]), "=r"((v)[7])                                \
                 : "r"(addr))
#define NFB_TMEM_LD16(addr, v)                                                           \
    asm volatile(                                                                        \
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                        \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"                \
        : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), \
          "=r"((v)[6]), "=r"((v)[7]), "=r"((v)[8]), "=r"((v)[9]), "=r"((v)[10]), "=r"((v)[11]), \
          "=r"((v)[12]), "=r"((v)[13]), "=r"((v)[14]), "=r"((v)[15])                     \
        : "r"(addr))
#define NFB_TMEM_LD32(addr, v)                                                           \
    asm volatile(                                                                        \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                        \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                        \
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"       \
        : "=r"((v)[0]), "=r"((v)[1]), "=r"((v)[2]), "=r"((v)[3]), "=r"((v)[4]), "=r"((v)[5]), \
          "=r"((v)[6]), "=r"((v)[7]), "=r"((v)[8]), "=r"((v)[9]), "=r"((v)[10]), "=r"((v)[11]), \
          "=r"((v)[12]), "=r"((v)[13]), "=r"((v)[14]), "=r"((v)[15]), "=r"((v)[16]),      \
          "=r"((v)[17]), "=r"((v)[18]), "=r"((v)[19]), "=r"((v)[20]), "=r"((v)[21]),      \
          "=r"((v)[22]), "=r"((v)[23]), "=r"((v)[24]), "=r"((v)[25]), "=r"((v)[26]),      \
          "=r"((v)[27]), "=r"((v)[28]), "=r"((v)[29]), "=r"((v)[30]), "=r"((v)[31])       \
        : "r"(addr))

// ----------------------------------------------------------------------------------------
// bf16 split helpers: v ~= hi + lo (+ lo2), each term a bf16 (round-to-nearest-even)
// ----------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_elem, float hi_elem) {
    uint32_t r;  // cvt.rn.bf16x2.f32 d, a, b: a -> upper half, b -> lower half
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo_elem, float hi_elem) {
    uint32_t r;  // cvt.rn.f16x2.f32 d, a, b: a -> upper half, b -> lower half
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_elem), "f"(lo_elem));
    return r;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t h) {  // .x = lower half, .y = upper half
    return __half22float2(*reinterpret_cast<const __half2*>(&h));
}
__device__ __forceinline__ float bf16_round(float v) {
    return __bfloat162float(__float2bfloat16_rn(v));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace nfb
