// Microbenchmark 2: does the ~630 clk per cp.async.bulk overlap across issuing threads / without waits?
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wseg_b200/csrc/ptx.cuh"
using namespace wsegk;

// mode 0: one thread, depth-2 ring (baseline)   mode 1: W warps, each its own depth-2 ring, concurrently
// mode 2: one thread issues `burst` copies back to back on one barrier, then waits (issue cost without the wait in between)
// mode 3: like 0 but lanes 0..burst-1 of ONE warp each issue a copy of bytes/burst in the same instruction
__global__ void __launch_bounds__(256, 1) k(const uint8_t* src, size_t span, uint32_t bytes, int mode, int W, int burst, int iters,
                                            long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[32];
    if (threadIdx.x == 0) {
        for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bar[i]), 1);
        mbar_fence_init();
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    long long t0 = clock64();
    if (mode == 0 || mode == 1) {
        if (lane == 0 && warp < W) {
            uint8_t* buf = smem + (size_t)warp * 2 * bytes;
            for (int i = 0; i < iters + 2; ++i) {
                const int s = i & 1;
                const uint32_t b = smem_u32(&bar[warp * 2 + s]);
                if (i >= 2) mbar_wait(b, ((i >> 1) - 1) & 1);
                if (i < iters) {
                    mbar_arrive_expect_tx(b, bytes);
                    tma_load_1d(smem_u32(buf + (size_t)s * bytes), src + ((size_t)(warp * 64 + i) * 65536) % span, bytes, b);
                }
            }
        }
    } else if (mode == 2) {
        if (threadIdx.x == 0) {
            for (int i = 0; i < iters / burst; ++i) {
                const uint32_t b = smem_u32(&bar[0]);
                mbar_arrive_expect_tx(b, bytes * burst);
                for (int j = 0; j < burst; ++j)
                    tma_load_1d(smem_u32(smem + (size_t)j * bytes), src + ((size_t)(i * burst + j) * 65536) % span, bytes, b);
                mbar_wait(b, i & 1);
            }
        }
    } else if (mode == 3) {
        if (warp == 0) {
            const uint32_t part = bytes / burst;
            for (int i = 0; i < iters + 2; ++i) {
                const int s = i & 1;
                const uint32_t b = smem_u32(&bar[s]);
                if (i >= 2) mbar_wait(b, ((i >> 1) - 1) & 1);
                if (i < iters) {
                    if (lane == 0) mbar_arrive_expect_tx(b, bytes);
                    __syncwarp();
                    if (lane < burst)
                        tma_load_1d(smem_u32(smem + (size_t)s * bytes + lane * part), src + ((size_t)i * 65536) % span + lane * part, part, b);
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = clock64() - t0;
}

int main() {
    const size_t span = 64ull << 20;
    uint8_t* src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
    long long* out; cudaMalloc(&out, 148 * 8);
    long long h[148];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    auto run = [&](const char* name, uint32_t bytes, int mode, int W, int burst, size_t smem, int copies) {
        const int iters = 64;
        for (int rep = 0; rep < 2; ++rep) { k<<<148, 256, smem>>>(src, span, bytes, mode, W, burst, iters, out); cudaDeviceSynchronize(); }
        cudaMemcpy(h, out, 148 * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < 148; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("%-46s bytes=%6u: %7lld clk, %5.0f clk per copy-op, %5.1f B/clk/SM\n", name, bytes, mx, (double)mx / copies,
               (double)copies * bytes / mx);
    };
    for (uint32_t bytes : {8192u, 21504u}) {
        run("1 thread, ring of 2", bytes, 0, 1, 1, 2 * bytes, 64);
        run("2 warps, each a ring of 2", bytes, 1, 2, 1, 4 * bytes, 128);
        run("4 warps, each a ring of 2", bytes, 1, 4, 1, 8 * bytes, 256);
        run("1 thread, 2 copies per barrier then wait", bytes, 2, 1, 2, 2 * bytes, 64);
        run("1 thread, 4 copies per barrier then wait", bytes, 2, 1, 4, 4 * bytes, 64);
        run("1 warp, 4 lanes issue 1/4 each (ring of 2)", bytes, 3, 1, 4, 2 * bytes, 64);
        run("1 warp, 8 lanes issue 1/8 each (ring of 2)", bytes, 3, 1, 8, 2 * bytes, 64);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
