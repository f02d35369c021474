// Issue cost vs execution time of tcgen05.mma.cta_group::2 (M = 256) for the shapes the PCM kernels use, of tcgen05.commit
// and of an mbarrier wait on a completed phase — what ONE issuing thread can sustain.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -cudart shared -I../../wseg_b200/csrc -o mma_issue mma_issue.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace wsegk;

struct Res { long long issue, commit, done; };

// kind: 0 = SS (A, B from shared memory), 1 = TS (A from TMEM)
__device__ void run_case(int kind, int N, int R, uint32_t tmem, uint32_t smem_a, uint32_t smem_b, uint32_t bar, uint32_t& phase, Res* out) {
    const uint32_t idesc = umma_idesc_f16(0, 256, N, 0, 0);
    const uint64_t da = umma_smem_desc(smem_a, 128 * 16, 128);
    const uint64_t db = umma_smem_desc(smem_b, (N / 2) * 16, 128);
    const long long t0 = clock64();
    for (int r = 0; r < R; ++r) {
        if (kind == 0) umma2_ss(tmem, da + static_cast<uint64_t>((r & 7) * 256), db + static_cast<uint64_t>((r & 7) * (N / 2) * 2), idesc, 1);
        else umma2_ts(tmem + 256, tmem + (r & 7) * 8, db + static_cast<uint64_t>((r & 7) * (N / 2) * 2), idesc, 1);
    }
    const long long t1 = clock64();
    umma2_commit(bar);
    const long long t2 = clock64();
    mbar_wait(bar, phase);
    phase ^= 1;
    const long long t3 = clock64();
    out->issue = t1 - t0; out->commit = t2 - t1; out->done = t3 - t0;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) bench(Res* res, long long* misc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar, bar2;
    __shared__ uint32_t tmem_base;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // 1.0 halves
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&bar2), 1); mbar_fence_init(); }
    if (warp == 1) { tmem_alloc_2cta(smem_u32(&tmem_base), 512); tmem_relinquish_2cta(); }
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            uint32_t phase = 0;
            const uint32_t a = smem_u32(smem), b = smem_u32(smem + 64 * 1024);
            int c = 0;
            const int Ns[6] = {64, 112, 128, 192, 224, 256};
            for (int rep = 0; rep < 2; ++rep) {           // second pass = warm
                c = 0;
                for (int i = 0; i < 6; ++i) run_case(0, Ns[i], 32, tmem, a, b, smem_u32(&bar), phase, res + c++);
                for (int i = 0; i < 4; ++i) run_case(1, Ns[i], 32, tmem, a, b, smem_u32(&bar), phase, res + c++);
                // a short burst (12 MMAs, as one S product) and one of 7 (one P V product)
                run_case(0, 112, 12, tmem, a, b, smem_u32(&bar), phase, res + c++);
                run_case(1, 64, 7, tmem, a, b, smem_u32(&bar), phase, res + c++);
                run_case(0, 224, 12, tmem, a, b, smem_u32(&bar), phase, res + c++);
            }
            // commit alone (nothing outstanding) and the wait right after it, x 8
            uint32_t ph2 = 0;
            long long tc = 0, tw = 0;
            for (int i = 0; i < 8; ++i) {
                const long long t0 = clock64();
                umma2_commit(smem_u32(&bar2));
                const long long t1 = clock64();
                mbar_wait(smem_u32(&bar2), ph2);
                ph2 ^= 1;
                const long long t2 = clock64();
                tc += t1 - t0; tw += t2 - t1;
            }
            misc[0] = tc / 8;
            misc[1] = tw / 8;
            long long t0, t1;
            // wait on a phase that completed long ago
            umma2_commit(smem_u32(&bar2));
            for (int i = 0; i < 2000; ++i) asm volatile("nanosleep.u32 20;");
            t0 = clock64();
            mbar_wait(smem_u32(&bar2), ph2);
            t1 = clock64();
            misc[2] = t1 - t0;
            t0 = clock64();
            t1 = clock64();
            misc[3] = t1 - t0;
        }
        __syncwarp();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) tmem_dealloc_2cta(tmem, 512);
}

int main() {
    Res* res; long long* misc;
    cudaMalloc(&res, 64 * sizeof(Res)); cudaMalloc(&misc, 64);
    cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    bench<<<2, 128, 200 * 1024>>>(res, misc);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    Res h[16]; long long m[4];
    cudaMemcpy(h, res, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(m, misc, sizeof(m), cudaMemcpyDeviceToHost);
    const char* names[13] = {"SS N=64 x32", "SS N=112 x32", "SS N=128 x32", "SS N=192 x32", "SS N=224 x32", "SS N=256 x32",
                             "TS N=64 x32", "TS N=112 x32", "TS N=128 x32", "TS N=192 x32", "SS N=112 x12", "TS N=64 x7", "SS N=224 x12"};
    const int cnt[13] = {32, 32, 32, 32, 32, 32, 32, 32, 32, 32, 12, 7, 12};
    const int nn[13] = {64, 112, 128, 192, 224, 256, 64, 112, 128, 192, 112, 64, 224};
    for (int i = 0; i < 13; ++i)
        printf("%-14s issue loop %6lld clk (%5.1f per MMA)  commit %4lld  issue->complete %6lld clk (%5.1f per MMA; ideal %5.1f at 8192 flop/clk/SM)\n", names[i],
               h[i].issue, double(h[i].issue) / cnt[i], h[i].commit, h[i].done, double(h[i].done) / cnt[i], 128.0 * nn[i] * 16 / 4096);
    printf("commit alone %lld clk; wait right after commit %lld clk; wait on an old phase %lld clk; clock64 pair %lld clk\n", m[0], m[1], m[2], m[3]);
    return 0;
}
