// Does a warp-synchronising instruction after tcgen05.commit stall the issuing warp until the committed MMAs complete?
// Pattern A: per iteration  if (elect) { 12 MMAs; commit; }  __syncwarp();     (the shape of the PCM issuer loops)
// Pattern B: the whole loop inside ONE elected branch.        Pattern C: as A but the branch is `lane == 0`, no __syncwarp.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -I../../wseg_b200/csrc -o mma_sync mma_sync.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace wsegk;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(int pattern, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[4];
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bars[i]), 1); mbar_fence_init(); }
    if (warp == 1) { tmem_alloc_2cta(smem_u32(&tmem_base), 512); tmem_relinquish_2cta(); }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    const uint32_t id_s = umma_idesc_f16(0, 256, 112, 0, 0);
    const uint64_t da = umma_smem_desc(a, 128 * 16, 128), db = umma_smem_desc(b, 56 * 16, 128);
    const int ITERS = 20;
    if (rank == 0 && warp == 0) {
        long long t0 = clock64(), t1 = 0;
        auto block = [&](int it) {
#pragma unroll
            for (int r = 0; r < 12; ++r)
                umma2_ss(tmem + (it % 3) * 128, da + static_cast<uint64_t>(r * 256), db + static_cast<uint64_t>(r * 112), id_s, r > 0);
            umma2_commit(smem_u32(&bars[0]));
        };
        if (pattern == 0) {
            for (int it = 0; it < ITERS; ++it) {
                if (elect_one()) block(it);
                __syncwarp();
            }
            t1 = clock64();
        } else if (pattern == 1) {
            if (elect_one()) {
                for (int it = 0; it < ITERS; ++it) block(it);
                t1 = clock64();
            }
            t1 = __shfl_sync(0xffffffffu, t1, __ffs(__ballot_sync(0xffffffffu, t1 != 0)) - 1);
        } else if (pattern == 2) {
            for (int it = 0; it < ITERS; ++it) {
                if (lane == 0) block(it);
            }
            t1 = clock64();
        } else if (pattern == 3) {                 // as A, plus a tcgen05.fence::before_thread_sync in front of the __syncwarp
            for (int it = 0; it < ITERS; ++it) {
                if (elect_one()) block(it);
                __syncwarp();
            }
            t1 = clock64();
        }
        if (lane == 0) {
            out[0] = (t1 - t0) / ITERS;
            umma2_commit(smem_u32(&bars[1]));
            mbar_wait(smem_u32(&bars[1]), 0);
            out[1] = (clock64() - t0) / ITERS;
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

int main() {
    long long* out;
    cudaMalloc(&out, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const char* names[] = {"A: if (elect) {12 MMA; commit}  __syncwarp()", "B: whole loop inside one elected branch", "C: if (lane == 0) {...}, no __syncwarp"};
    for (int p = 0; p < 3; ++p) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaMemset(out, 0, 64);
            bench<<<2, 128, 200 * 1024>>>(p, out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2];
            cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
            if (rep) printf("%s %-48s issue loop %5lld clk/iter, all complete %5lld clk/iter (pipe: 12 x 64 = 768)\n", cudaGetErrorString(e), names[p], h[0], h[1]);
        }
    }
    return 0;
}
