// Microbenchmark: per-SM streaming bandwidth L2 -> shared memory.
//   mode 0: cp.async.bulk (UBLKCP) records into a ring, consumer = mbarrier wait only
//   mode 1: cp.async (LDGSTS) 16 B/thread from W producer warps, commit groups
#include <cstdio>
#include <cuda_runtime.h>
#include "../normalizing-flows_b200/csrc/nfb_common.cuh"
void nfb_set_error(const char*, ...) {}
using namespace nfb;

__global__ void __launch_bounds__(64, 1) bulk_stream(const uint8_t* src, size_t src_bytes, int rec_bytes,
                                                     int slots, int nrec, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    const uint32_t sbase = smem_u32(smem);
    if (threadIdx.x == 0) {
        for (int i = 0; i < slots; ++i) { mbar_init(smem_u32(&full[i]), 1); mbar_init(smem_u32(&empty[i]), 1); }
        fence_mbar_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    long long t0 = clock64();
    if (warp == 0) {
        uint32_t slot = 0, par = 0;
        size_t off = ((size_t)blockIdx.x * 7919 * 1024) % (src_bytes - rec_bytes);
        off &= ~(size_t)1023;
        for (int r = 0; r < nrec; ++r) {
            mbar_wait(smem_u32(&empty[slot]), par ^ 1, nullptr, 0);
            if (elect_one_sync()) {
                mbar_expect_tx(smem_u32(&full[slot]), rec_bytes);
                bulk_g2s(sbase + slot * rec_bytes, src + off, rec_bytes, smem_u32(&full[slot]));
            }
            __syncwarp();
            off += rec_bytes;
            if (off + rec_bytes > src_bytes) off = 0;
            if (++slot == (uint32_t)slots) { slot = 0; par ^= 1; }
        }
    } else {
        uint32_t slot = 0, par = 0;
        for (int r = 0; r < nrec; ++r) {
            mbar_wait(smem_u32(&full[slot]), par, nullptr, 0);
            if (elect_one_sync()) mbar_arrive(smem_u32(&empty[slot]));
            __syncwarp();
            if (++slot == (uint32_t)slots) { slot = 0; par ^= 1; }
        }
        if (threadIdx.x == 32) out[blockIdx.x] = clock64() - t0;
    }
}

// LDGSTS: `nw` warps each copy 512 B per instruction; groups of `depth` instructions per commit
__global__ void __launch_bounds__(256, 1) ldgsts_stream(const uint8_t* src, size_t src_bytes, int total_bytes,
                                                        int inflight_groups, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const int nthreads = blockDim.x;
    size_t off = ((size_t)blockIdx.x * 7919 * 1024) % (src_bytes / 2);
    off &= ~(size_t)1023;
    const int chunk = nthreads * 16;  // bytes per block-wide instruction
    const int ring = 65536;
    long long t0 = clock64();
    int issued = 0, pos = 0, g = 0;
    while (issued < total_bytes) {
        // one commit group = 4 instructions (4 * chunk bytes)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t dst = sbase + ((pos + threadIdx.x * 16) & (ring - 1));
            const uint8_t* s = src + ((off + issued + threadIdx.x * 16) % (src_bytes - 16));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(s) : "memory");
            pos += chunk; issued += chunk;
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        ++g;
        if (g >= inflight_groups) {
            if (inflight_groups == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
            else if (inflight_groups == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else if (inflight_groups == 4) asm volatile("cp.async.wait_group 3;" ::: "memory");
            else asm volatile("cp.async.wait_group 7;" ::: "memory");
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t src_bytes = 8u << 20;  // 8 MB: L2 resident
    uint8_t* src; long long* d;
    cudaMalloc(&src, src_bytes); cudaMemset(src, 1, src_bytes);
    cudaMalloc(&d, 8 * 256);
    long long h[256];
    for (int grid : {1, 148}) {
        for (int rec : {4096, 8192, 16384, 32768}) for (int slots : {1, 2, 4, 8}) {
            if ((size_t)rec * slots > 200 * 1024) continue;
            const int nrec = (4 << 20) / rec;
            cudaFuncSetAttribute(bulk_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, rec * slots);
            for (int rep = 0; rep < 2; ++rep) bulk_stream<<<grid, 64, rec * slots>>>(src, src_bytes, rec, slots, nrec, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("bulk  grid=%3d rec=%5d slots=%d: %6.1f B/clk/SM  (%7.0f cyc/rec) %s\n", grid, rec, slots,
                   (double)rec * nrec / mx, (double)mx / nrec, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
        for (int nt : {32, 64, 128, 256}) for (int infl : {1, 2, 4, 8}) {
            const int total = 4 << 20;
            cudaFuncSetAttribute(ldgsts_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
            for (int rep = 0; rep < 2; ++rep) ldgsts_stream<<<grid, nt, 65536>>>(src, src_bytes, total, infl, d);
            cudaError_t e = cudaDeviceSynchronize();
            cudaMemcpy(h, d, 8 * grid, cudaMemcpyDeviceToHost);
            long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("ldgsts grid=%3d threads=%3d groups_in_flight=%d (%5d B each): %6.1f B/clk/SM %s\n", grid, nt, infl,
                   nt * 64, (double)total / mx, e == cudaSuccess ? "" : cudaGetErrorString(e));
        }
    }
    return 0;
}
