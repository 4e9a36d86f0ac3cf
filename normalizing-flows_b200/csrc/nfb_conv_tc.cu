// nfb_conv_tc.cu -- stride-1 "same" convolution of the Glow conditioner (nets/cnn.py:33-61, ConvNet2d) as an
// implicit GEMM on the sm_100a tensor core.
//
//   y[b, n, h, w] = act( sum_{c,kh,kw} W[n, c, kh, kw] x[b, c0+c, h+kh-p, w+kw-p] + bias[n] )
//   M = B*H*W pixels (128 per CTA = the 128 TMEM lanes), N = cout (<= 256, one accumulator), K in chunks of 64.
//   K order: k = cb*(T*16) + tap*16 + ci for channel c = 16 cb + ci and tap = kh*k + kw (T = k*k), i.e. blocks
//   of 16 channels, tap-major inside a block.  A 16-wide group of k is then ONE tap of 16 consecutive channels:
//   one bounds check and one base address per group, the 16 loads differ by the plane stride only (3 instead of
//   ~20 instructions per gathered element), while all T taps of a channel block stay within 2-3 chunks, so the
//   shifted re-reads of the same 16 planes hit L1.  Channels are zero-padded to a multiple of 16.
//
// Numerics: the same split-bf16 scheme as the spline conditioner (a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo,
// fp32 accumulation in TMEM, weights packed with the accumulate-truncation gain, see nfb_api.cu kAccStepGain).
//
// Roles (576 threads, persistent: CTA b takes 128-pixel tiles b, b+grid, ...):
//   warps 0-15 builders, 4 per SM sub-partition: thread = (pixel, quarter of the K-chunk).  Per K-chunk it gathers
//              16 im2col values of its pixel (consecutive lanes = consecutive pixels -> coalesced; the taps of a
//              channel hit L1) -- the loads of chunk k+1 are issued before chunk k is converted -- splits them to
//              bf16 hi/lo and stores two 16-byte chunks per tile straight into the SWIZZLE_128B A tiles (3-6
//              stages, as many as the weight ring leaves room for).  The epilogue of tile i-1 (TMEM -> bias ->
//              LeakyReLU -> NCHW store, coalesced per channel) runs AFTER tile i has been built, so it overlaps
//              tile i's MMAs; the accumulator is double-buffered in TMEM.
//   warp 16    one elected lane streams the packed weight records ([n_pad x 64] hi | lo, pre-swizzled) with
//              1-D bulk TMA into a 2-slot ring.
//   warp 17    one elected lane issues 12 tcgen05.mma (M=128, N=n_pad, K=16) per chunk; tcgen05.commit frees the
//              A stage and the weight slot; owns the TMEM allocation.
#include "nfb_kernels.h"

namespace nfb {

namespace {
constexpr int kCtBuildWarps = 16;
constexpr int kCtThreads = 32 * kCtBuildWarps + 64;
constexpr int kCtMaxStages = 6;
constexpr uint32_t kCtTileA = 16384;   // [128 x 64] bf16
constexpr uint32_t kCtStage = 2 * kCtTileA;  // hi | lo
constexpr uint32_t kCtBarBytes = 32 * 8;
constexpr uint32_t kCtSmemMax = 232448;
constexpr int kCtMaxSlots = 8;
enum { CB_AFULL = 0, CB_AEMPTY = 6, CB_WFULL = 12, CB_WEMPTY = 20, CB_ACCFULL = 28, CB_ACCEMPTY = 30 };

__device__ __forceinline__ uint32_t ct_chunk_off(int r, int c8) {
    return (r >> 3) * 1024 + (r & 7) * 128 + ((c8 ^ (r & 7)) << 4);
}
__device__ __forceinline__ void ct_st_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
}  // namespace

struct ConvTcParams {
    const float* x; float* y; const float* bias; const uint8_t* wstream;
    long long M; int ctot, c0, cin, H, W, cout, ks, n_pad, k_chunks, stages, w_slots; uint32_t slot_bytes; float leaky; int* err;
};

// shared memory: [stages x (A hi | A lo)] [w_slots x weight slot] [barriers] [tmem ptr]
__global__ void __launch_bounds__(kCtThreads, 1) conv_tc_kernel(const ConvTcParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int S = p.stages;
    const uint32_t offW = (uint32_t)S * kCtStage;
    const int WS = p.w_slots;
    const uint32_t offBars = offW + (uint32_t)WS * p.slot_bytes;
    const uint32_t bars = sbase + offBars;
    auto bar = [bars](int i) { return bars + 8u * i; };
    const uint32_t tcols = p.n_pad <= 32 ? 32u : p.n_pad <= 64 ? 64u : p.n_pad <= 128 ? 128u : 256u;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kCtMaxStages; ++i) {
            mbar_init(bar(CB_AFULL + i), kCtBuildWarps);  // one arrive per builder warp
            mbar_init(bar(CB_AEMPTY + i), 1);             // tcgen05.commit
        }
        for (int i = 0; i < kCtMaxSlots; ++i) {
            mbar_init(bar(CB_WFULL + i), 1);              // expect_tx
            mbar_init(bar(CB_WEMPTY + i), 1);             // tcgen05.commit
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(bar(CB_ACCFULL + i), 1);            // tcgen05.commit
            mbar_init(bar(CB_ACCEMPTY + i), kCtBuildWarps);
        }
        fence_mbar_init();
    }
    if (warp == kCtBuildWarps + 1) {
        tmem_alloc(sbase + offBars + kCtBarBytes, 2 * tcols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + offBars + kCtBarBytes);
    const int KC = p.k_chunks;
    const uint32_t rec_bytes = (uint32_t)p.n_pad * 256u;  // hi + lo
    const long long n_tiles = (p.M + 127) / 128;

    if (warp == kCtBuildWarps) {
        // ------------------------------ weight producer ------------------------------------
        uint32_t ws = 0, ws_use = 0;  // ring position over all tiles of this CTA
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x)
            for (int kc = 0; kc < KC; ++kc) {
                if (ws_use > 0) mbar_wait(bar(CB_WEMPTY + ws), (ws_use - 1) & 1u, p.err, 700 + ws);
                if (elect_one_sync()) {
                    mbar_expect_tx(bar(CB_WFULL + ws), rec_bytes);
                    bulk_g2s(sbase + offW + ws * p.slot_bytes, p.wstream + (size_t)kc * rec_bytes, rec_bytes,
                             bar(CB_WFULL + ws));
                }
                __syncwarp();
                if (++ws == (uint32_t)WS) { ws = 0; ++ws_use; }
            }
    } else if (warp == kCtBuildWarps + 1) {
        // ------------------------------ MMA issuer ------------------------------------------
        const uint32_t idesc = umma_idesc_bf16(128, (uint32_t)p.n_pad);
        uint32_t ws = 0, ws_use = 0, st = 0, st_use = 0, it = 0;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const uint32_t ab = it & 1u;  // accumulator buffer
            if (it >= 2) mbar_wait(bar(CB_ACCEMPTY + ab), ((it >> 1) - 1) & 1u, p.err, 750 + ab);
            const uint32_t d = tmem + ab * tcols;
            for (int kc = 0; kc < KC; ++kc) {
                mbar_wait(bar(CB_AFULL + st), st_use & 1u, p.err, 710 + st);
                mbar_wait(bar(CB_WFULL + ws), ws_use & 1u, p.err, 720 + ws);
                tc_fence_after();
                if (elect_one_sync()) {
                    const uint64_t a_hi = umma_desc_sw128(sbase + st * kCtStage);
                    const uint64_t a_lo = umma_desc_sw128(sbase + st * kCtStage + kCtTileA);
                    const uint64_t w_hi = umma_desc_sw128(sbase + offW + ws * p.slot_bytes);
                    const uint64_t w_lo = umma_desc_sw128(sbase + offW + ws * p.slot_bytes + (uint32_t)p.n_pad * 128u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_bf16(d, a_hi + 2 * j, w_hi + 2 * j, idesc, (kc | j) ? 1u : 0u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_bf16(d, a_lo + 2 * j, w_hi + 2 * j, idesc, 1u);
#pragma unroll
                    for (int j = 0; j < 4; ++j) umma_bf16(d, a_hi + 2 * j, w_lo + 2 * j, idesc, 1u);
                    umma_commit(bar(CB_AEMPTY + st));
                    umma_commit(bar(CB_WEMPTY + ws));
                    if (kc == KC - 1) umma_commit(bar(CB_ACCFULL + ab));
                }
                __syncwarp();
                if (++st == (uint32_t)S) { st = 0; ++st_use; }
                if (++ws == (uint32_t)WS) { ws = 0; ++ws_use; }
            }
        }
    } else {
        // ------------------------------ im2col builders / epilogue -------------------------
        const int q = warp & 3, wh = warp >> 2;  // TMEM lane quadrant, K-quarter / column group
        const int r = q * 32 + lane;             // tile row = TMEM lane
        const int HW = p.H * p.W, pad = p.ks >> 1, ks = p.ks, kk2 = ks * ks;
        const uint32_t tlane = tmem + ((uint32_t)(q * 32) << 16);
        uint32_t st = 0, st_use = 0, it = 0;

        auto epilogue = [&](long long tile, uint32_t i) {
            const uint32_t ab = i & 1u;
            mbar_wait(bar(CB_ACCFULL + ab), (i >> 1) & 1u, p.err, 740 + ab);
            tc_fence_after();
            const long long m = tile * 128 + r;
            const bool live = m < p.M;
            const long long bi = live ? m / HW : 0;
            const int pix = live ? (int)(m - bi * HW) : 0;
            float* yb = p.y + bi * (long long)p.cout * HW + pix;
            for (int n0 = wh * 16; n0 < p.n_pad; n0 += 64) {
                uint32_t acc[16];
                NFB_TMEM_LD16(tlane + ab * tcols + n0, acc);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int n = n0 + j;
                    if (live && n < p.cout) {
                        float v = __uint_as_float(acc[j]) + (p.bias ? __ldg(p.bias + n) : 0.f);
                        if (p.leaky >= 0.f) v = v >= 0.f ? v : v * p.leaky;
                        yb[(long long)n * HW] = v;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(CB_ACCEMPTY + ab));
        };

        long long prev_tile = -1;
        for (long long t = blockIdx.x; t < n_tiles; t += gridDim.x, ++it) {
            const long long m = t * 128 + r;
            const bool live = m < p.M;
            const long long bi = live ? m / HW : 0;
            const int pix = live ? (int)(m - bi * HW) : 0;
            const int h = pix / p.W, w = pix - h * p.W;
            const float* xb = p.x + (bi * p.ctot + p.c0) * (long long)HW;
            // this thread's 16 k indices of chunk kc = group G = 4 kc + wh: one tap of channel block cb.  The
            // calls come in increasing kc, so (cb, tap) is carried along instead of divided out each time, and
            // the 16 loads use one 32-bit offset stepped by the plane stride.
            int g_cb = ks == 1 ? wh : wh / kk2, g_tap = ks == 1 ? 0 : wh % kk2;
            const int inv_ks = 65536 / ks + 1;  // tap / ks for tap < 25
            auto gather = [&](float (&v)[16]) {
                const int kh = (g_tap * inv_ks) >> 16, kw = g_tap - kh * ks;
                const int hh = h + kh - pad, ww = w + kw - pad;
                const bool ok = live && hh >= 0 && hh < p.H && ww >= 0 && ww < p.W;
                const int cbase = g_cb * 16;
                const int nvalid = ok ? p.cin - cbase : 0;  // channels of this block that exist (<= 0: none)
                const float* src = xb + ((cbase * p.H + hh) * p.W + ww);  // one pointer, stepped by the plane stride
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    v[j] = 0.f;
                    if (j < nvalid) v[j] = __ldg(src);
                    src += HW;
                }
                if (ks == 1) g_cb += 4;
                else { g_tap += 4; while (g_tap >= kk2) { g_tap -= kk2; ++g_cb; } }
            };
            auto emit = [&](const float (&v)[16]) {  // split, store into the next free A stage, publish
                if (st_use > 0) mbar_wait(bar(CB_AEMPTY + st), (st_use - 1) & 1u, p.err, 730 + st);
                const uint32_t t_hi = sbase + st * kCtStage, t_lo = t_hi + kCtTileA;
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float a = v[8 * g + 2 * i], b = v[8 * g + 2 * i + 1];
                        hi[i] = pack_bf16x2(a, b);
                        lo[i] = pack_bf16x2(a - __uint_as_float(hi[i] << 16), b - __uint_as_float(hi[i] & 0xffff0000u));
                    }
                    const uint32_t off = ct_chunk_off(r, wh * 2 + g);
                    ct_st_v4(t_hi + off, hi[0], hi[1], hi[2], hi[3]);
                    ct_st_v4(t_lo + off, lo[0], lo[1], lo[2], lo[3]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar(CB_AFULL + st));
                if (++st == (uint32_t)S) { st = 0; ++st_use; }
            };
            // three register sets in rotation: the loads of chunks k+1 and k+2 are in flight while chunk k is
            // converted (the gathers come from HBM/L2 at ~2 k cycles; one chunk of work is ~0.6 k)
            float a0[16], a1[16], a2[16];
            gather(a0);
            if (KC > 1) gather(a1);
            for (int kc = 0; kc < KC; kc += 3) {
                if (kc + 2 < KC) gather(a2);
                emit(a0);
                if (kc + 1 < KC) {
                    if (kc + 3 < KC) gather(a0);
                    emit(a1);
                }
                if (kc + 2 < KC) {
                    if (kc + 4 < KC) gather(a1);
                    emit(a2);
                }
            }
            if (prev_tile >= 0) epilogue(prev_tile, it - 1);  // overlaps this tile's MMAs
            prev_tile = t;
        }
        if (prev_tile >= 0) epilogue(prev_tile, it - 1);
    }
    __syncthreads();
    if (warp == kCtBuildWarps + 1) tmem_dealloc(tmem, 2 * tcols);
}

// weights [cout, cin, k, k] fp32 -> per K-chunk record [n_pad x 64] hi | lo, SWIZZLE_128B, K in the kernel's order
__global__ void conv_pack_kernel(const float* __restrict__ w, int cout, int cin, int T, int n_pad, int k_chunks,
                                 float gain, uint8_t* __restrict__ out) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)k_chunks * n_pad * 64;
    if (idx >= total) return;
    const int kk = (int)(idx & 63);
    const int n = (int)((idx >> 6) % n_pad);
    const int kc = (int)(idx / ((long long)n_pad * 64));
    const int G = kc * 4 + (kk >> 4), ci = kk & 15;
    const int cb = G / T, tap = G - cb * T, c = cb * 16 + ci;
    const float v = (n < cout && c < cin) ? w[((long long)n * cin + c) * T + tap] * gain : 0.f;
    const size_t off = (size_t)(n >> 3) * 1024 + (n & 7) * 128 + (((kk >> 3) ^ (n & 7)) << 4) + (kk & 7) * 2;
    uint8_t* rec = out + (size_t)kc * n_pad * 256;
    const __nv_bfloat16 hi = __float2bfloat16_rn(v);
    *reinterpret_cast<__nv_bfloat16*>(rec + off) = hi;
    *reinterpret_cast<__nv_bfloat16*>(rec + (size_t)n_pad * 128 + off) = __float2bfloat16_rn(v - __bfloat162float(hi));
}

bool conv_tc_supported(int cin, int cout, int ks) {
    const int K = cin * ks * ks;
    // small square maps (the folded ActNorm + Invertible1x1Conv, which transforms z itself) stay on the fp32 kernel
    return (ks == 1 || ks == 3 || ks == 5) && cout >= 1 && cout <= 256 && (cout > 64 || K > 64);
}

int launch_conv2d_tc(const float* x, int ctot, int c0, const float* w, const float* bias, float* y, long long B,
                     int cin, int H, int W, int cout, int ks, float leaky, float gain_per_step, int* err,
                     cudaStream_t st) {
    static PerDevice per_dev;  // attribute + SM count of the device this launch goes to (not the first one seen)
    const int sm_count = per_dev.ensure([] {
        return cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCtSmemMax);
    });
    if (sm_count < 0) return NFB_ERR_CUDA;
    const long long M = B * H * W;
    if (M == 0) return NFB_OK;
    const int T = ks * ks;
    const int groups = (cin + 15) / 16 * T;  // 16-wide k groups: (channel block, tap)
    const int k_chunks = (groups + 3) / 4;
    const int n_pad = (cout + 15) / 16 * 16;
    const size_t bytes = (size_t)k_chunks * n_pad * 256;
    void* scratch = nullptr;
    NFB_CUDA(cudaMallocAsync(&scratch, bytes, st));  // stream-ordered: freed after the kernel that reads it
    const long long total = (long long)k_chunks * n_pad * 64;
    const float gain = 1.f + gain_per_step * (float)(3 * 4 * k_chunks);
    conv_pack_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, cout, cin, T, n_pad, k_chunks, gain,
                                                                      static_cast<uint8_t*>(scratch));
    ConvTcParams p{};
    p.x = x; p.y = y; p.bias = bias; p.wstream = static_cast<const uint8_t*>(scratch);
    p.M = M; p.ctot = ctot; p.c0 = c0; p.cin = cin; p.H = H; p.W = W; p.cout = cout; p.ks = ks;
    p.n_pad = n_pad; p.k_chunks = k_chunks; p.leaky = leaky; p.err = err;
    p.slot_bytes = ((uint32_t)n_pad * 256u + 1023u) & ~1023u;
    // weight ring: up to 128 KB / 8 slots (small-N records are latency-, not bandwidth-limited); A stages: the rest
    int w_slots = (int)(131072u / p.slot_bytes);
    w_slots = w_slots < 2 ? 2 : (w_slots > kCtMaxSlots ? kCtMaxSlots : w_slots);
    const uint32_t fixed = (uint32_t)w_slots * p.slot_bytes + kCtBarBytes + 16;
    int stages = (int)((kCtSmemMax - fixed) / kCtStage);
    stages = stages > kCtMaxStages ? kCtMaxStages : stages;  // 3 at N = 256, 6 at N <= 64
    p.stages = stages;
    p.w_slots = w_slots;
    const uint32_t smem = (uint32_t)stages * kCtStage + fixed;
    const long long n_tiles = (M + 127) / 128;
    const unsigned grid = (unsigned)(n_tiles < sm_count ? n_tiles : sm_count);
    conv_tc_kernel<<<grid, kCtThreads, smem, st>>>(p);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) {
        nfb_set_error("conv_tc launch: %s", cudaGetErrorString(e));
        return NFB_ERR_CUDA;
    }
    return NFB_OK;
}

}  // namespace nfb
