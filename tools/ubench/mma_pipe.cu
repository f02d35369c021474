// Steady-state issue patterns of the PCM forward on one CTA pair, without data dependencies: how long does an iteration of
// [12 SS MMAs N=112 | commits | 7 TS MMAs N=64 | commits] take when issued (a) by one thread, (b) by two warps, (c) with
// mbarrier waits on completed phases in between, (d) with 8 ALU-saturating warps sharing the sub-partitions.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -I../../wseg_b200/csrc -o mma_pipe mma_pipe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace wsegk;

__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}

struct Cfg { int mode; int commits_s; int commits_pv; int waits; int alu; int test_wait; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(384, 1) bench(Cfg cfg, long long* out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bars[8];
    __shared__ __align__(8) uint64_t old_bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int stop;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 8; ++i) mbar_init(smem_u32(&bars[i]), 1);
        mbar_init(smem_u32(&old_bar), 1);
        mbar_fence_init();
        stop = 0;
    }
    if (warp == 1) { tmem_alloc_2cta(smem_u32(&tmem_base), 512); tmem_relinquish_2cta(); }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
    const uint32_t id_s = umma_idesc_f16(0, 256, 112, 0, 0), id_o = umma_idesc_f16(0, 256, 64, 0, 0);
    const uint64_t da = umma_smem_desc(a, 128 * 16, 128), db = umma_smem_desc(b, 56 * 16, 128), dc = umma_smem_desc(b, 32 * 16, 128);
    const int ITERS = 40;
    auto wait_old = [&]() {
        if (cfg.test_wait) { while (!mbar_test_wait(smem_u32(&old_bar), 0)) {} }
        else mbar_wait(smem_u32(&old_bar), 0);
    };
    auto issue_s = [&](int it) {
        for (int w = 0; w < cfg.waits; ++w) wait_old();
        for (int r = 0; r < 12; ++r)
            umma2_ss(tmem + (it % 3) * 128, da + static_cast<uint64_t>(r * 256), db + static_cast<uint64_t>(r * 112), id_s, r > 0);
        for (int c = 0; c < cfg.commits_s; ++c) umma2_commit(smem_u32(&bars[c]));
    };
    auto issue_pv = [&](int it) {
        for (int w = 0; w < cfg.waits; ++w) wait_old();
        for (int r = 0; r < 7; ++r)
            umma2_ts(tmem + 384, tmem + (it % 3) * 128 + r * 8, dc + static_cast<uint64_t>(r * 64), id_o, 1);
        for (int c = 0; c < cfg.commits_pv; ++c) umma2_commit(smem_u32(&bars[4 + c]));
    };
    if (rank == 0 && warp == 0 && elect_one()) {
        umma2_commit(smem_u32(&old_bar));            // completes at once: the "old phase" every wait polls
    }
    __syncthreads();
    if (rank == 0 && (warp == 1 || warp == 3)) {
        if (elect_one()) {
            const long long t0 = clock64();
            if (cfg.mode == 0 && warp == 1) {
                for (int it = 0; it < ITERS; ++it) { issue_pv(it); issue_s(it); }
            } else if (cfg.mode == 1) {
                if (warp == 1) for (int it = 0; it < ITERS; ++it) issue_s(it);
                else for (int it = 0; it < ITERS; ++it) issue_pv(it);
            }
            const long long t1 = clock64();
            // drain: one more commit on an own barrier and wait for it
            const uint32_t fin = smem_u32(&bars[warp == 1 ? 3 : 7]);
            umma2_commit(fin);
            // count completed phases of `fin` is unknown when commits_* used it: it is reserved (indices 3 and 7 unused above as commits <= 3)
            mbar_wait(fin, 0);
            const long long t2 = clock64();
            out[(warp == 1 ? 0 : 2)] = (t1 - t0) / ITERS;
            out[(warp == 1 ? 0 : 2) + 1] = (t2 - t0) / ITERS;
        }
        __syncwarp();
        if (warp == 1) stop = 1;
    } else if (warp >= 4 && cfg.alu) {
        // ALU-saturating filler on every sub-partition (2 warps each), as the ReLU / pack warps are
        uint32_t x = threadIdx.x, y = 0x3c003c00u;
        int n = 0;
        while (!stop && n < 2000000) {
#pragma unroll
            for (int i = 0; i < 64; ++i) { x = (x & y) ^ (x >> 3); y = y + (x | 1u); }
            ++n;
        }
        sink[threadIdx.x] = __uint_as_float(x ^ y);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

int main() {
    long long* out; float* sink;
    cudaMalloc(&out, 64); cudaMalloc(&sink, 4096);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const Cfg cfgs[] = {
        {0, 0, 0, 0, 0, 0}, {0, 1, 1, 0, 0, 0}, {0, 3, 2, 0, 0, 0}, {0, 3, 2, 1, 0, 0}, {0, 3, 2, 2, 0, 0}, {0, 3, 2, 2, 0, 1}, {0, 3, 2, 2, 1, 0}, {0, 1, 1, 0, 1, 0},
        {1, 1, 1, 0, 0, 0}, {1, 2, 1, 0, 0, 0}, {1, 2, 1, 2, 0, 0}, {1, 2, 1, 2, 0, 1}, {1, 2, 1, 2, 1, 0}, {1, 1, 1, 1, 1, 1},
    };
    printf("pipe time per iteration at the measured rates: 12 x 64 + 7 x 36 = 1020 clk\n");
    for (const Cfg& c : cfgs) {
        cudaMemset(out, 0, 64);
        bench<<<2, 384, 200 * 1024>>>(c, out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[4];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%s issuers=%d commits S/PV=%d/%d waits/block=%d (%s) alu-filler=%d : issue loop %5lld clk/iter, complete %5lld clk/iter", cudaGetErrorString(e),
               c.mode + 1, c.commits_s, c.commits_pv, c.waits, c.test_wait ? "test_wait" : "try_wait", c.alu, h[0], h[1]);
        if (c.mode == 1) printf("  | PV warp: %5lld / %5lld", h[2], h[3]);
        printf("\n");
    }
    return 0;
}
