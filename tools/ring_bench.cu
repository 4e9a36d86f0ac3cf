// Microbenchmark: bulk-copy ring consumed by tcgen05.mma with tcgen05.commit slot release
#include <cstdio>
#include <cuda_runtime.h>
#include "../normalizing-flows_b200/csrc/nfb_common.cuh"
void nfb_set_error(const char*, ...) {}
using namespace nfb;

template <bool TS>
__global__ void __launch_bounds__(64, 1) ring(const uint8_t* src, size_t src_bytes, int rec_bytes, int slots,
                                              int nrec, int mma_per_rec, int n, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16], empty[16], done;
    __shared__ uint32_t tptr;
    const uint32_t sbase = smem_u32(smem);
    const uint32_t abase = sbase;                 // 16 KB A tile
    const uint32_t wbase = sbase + 16384;
    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        mbar_init(smem_u32(&done), 1);
        fence_mbar_init();
    }
    for (int i = threadIdx.x; i < 4096; i += 64) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x < 32) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tptr;
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    if (warp == 0) {
        uint32_t slot = 0, par = 0;
        size_t off = 0;
        for (int r = 0; r < nrec; ++r) {
            mbar_wait(smem_u32(&empty[slot]), par ^ 1, nullptr, 0);
            if (elect_one_sync()) {
                mbar_expect_tx(smem_u32(&full[slot]), rec_bytes);
                bulk_g2s(wbase + slot * rec_bytes, src + off, rec_bytes, smem_u32(&full[slot]));
            }
            __syncwarp();
            off += rec_bytes;
            if (off + rec_bytes > src_bytes) off = 0;
            if (++slot == (uint32_t)slots) { slot = 0; par ^= 1; }
        }
    } else {
        uint32_t slot = 0, par = 0;
        const uint64_t ad = umma_desc_sw128(abase);
        const uint32_t idesc = umma_idesc_bf16(128, n);
        for (int r = 0; r < nrec; ++r) {
            mbar_wait(smem_u32(&full[slot]), par, nullptr, 0);
            tc_fence_after();
            if (elect_one_sync()) {
                const uint64_t bd = umma_desc_sw128(wbase + slot * rec_bytes);
                for (int m = 0; m < mma_per_rec; ++m) {
                    if (TS) umma_bf16_ts(tmem, tmem + 256 + (m & 3) * 8, bd + 2 * (m & 3), idesc, 1u);
                    else umma_bf16(tmem, ad + 2 * (m & 3), bd + 2 * (m & 3), idesc, 1u);
                }
                if (mma_per_rec) umma_commit(smem_u32(&empty[slot]));
                else mbar_arrive(smem_u32(&empty[slot]));
            }
            __syncwarp();
            if (++slot == (uint32_t)slots) { slot = 0; par ^= 1; }
        }
        if (elect_one_sync()) umma_commit(smem_u32(&done));
        __syncwarp();
        mbar_wait(smem_u32(&done), 0, nullptr, 0);
        if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    const size_t src_bytes = 4u << 20;
    uint8_t* src; long long* d;
    cudaMalloc(&src, src_bytes); cudaMemset(src, 0, src_bytes);
    cudaMalloc(&d, 8 * 256);
    long long h[256];
    const int grid = 148;
    for (int ts = 0; ts < 2; ++ts)
    for (int rec : {12288, 16384, 32768}) for (int slots : {2, 4, 6}) for (int mpr : {0, 4, 8, 12}) {
        const int n = rec / 128 > 256 ? 256 : rec / 128;
        if ((size_t)rec * slots + 16384 > 220 * 1024) continue;
        const int nrec = 512;
        const size_t sm = 16384 + (size_t)rec * slots;
        auto k = ts ? ring<true> : ring<false>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        for (int rep = 0; rep < 2; ++rep) k<<<grid, 64, sm>>>(src, src_bytes, rec, slots, nrec, mpr, n, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%s rec=%5d (N=%3d) slots=%d mma/rec=%2d: %7.0f cyc/rec  (mma ideal %5.0f)  %5.1f B/clk %s\n", ts ? "TS" : "SS",
               rec, n, slots, mpr, (double)mx / nrec, mpr * n / 2.0, (double)rec * nrec / mx,
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    return 0;
}
