// Microbenchmark: per-SM throughput of cp.async.bulk (1-D, UBLKCP) from L2-resident global memory into shared memory.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulkcp bulkcp.cu && ./bulkcp
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../wseg_b200/csrc/ptx.cuh"
using namespace wsegk;

// each CTA: `depth` buffers of `bytes`; loads `iters` copies round-robin, each from a different address
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t span, uint32_t bytes, int depth, int iters, int shared_src,
                                            long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) mbar_init(smem_u32(&bar[i]), 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const size_t base = shared_src ? 0 : (size_t)blockIdx.x * 7919 * 1024 % span;
        long long t0 = clock64();
        for (int i = 0; i < iters + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) mbar_wait(smem_u32(&bar[s]), ((i / depth) - 1) & 1);
            if (i < iters) {
                mbar_arrive_expect_tx(smem_u32(&bar[s]), bytes);
                const size_t off = (base + (size_t)i * bytes) % (span - bytes);
                tma_load_1d(smem_u32(smem + (size_t)s * bytes), src + (off & ~(size_t)127), bytes, smem_u32(&bar[s]));
            }
        }
        long long t1 = clock64();
        out[blockIdx.x] = t1 - t0;
    }
}

int main() {
    const size_t span = 64ull << 20;   // 64 MB, L2 resident after the first pass
    uint8_t* src; cudaMalloc(&src, span); cudaMemset(src, 1, span);
    long long* out; cudaMalloc(&out, 148 * 8);
    long long h[148];
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const uint32_t sizes[] = {2048, 8192, 21504, 43008};
    for (int shared_src = 0; shared_src < 2; ++shared_src)
    for (int grid : {1, 148})
    for (uint32_t bytes : sizes)
    for (int depth : {1, 2, 4, 8}) {
        if ((size_t)depth * bytes > 190 * 1024) continue;
        const int iters = 64;
        for (int rep = 0; rep < 2; ++rep) {
            k<<<grid, 128, (size_t)depth * bytes>>>(src, span, bytes, depth, iters, shared_src, out);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, out, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("same_src=%d grid=%3d bytes=%6u depth=%d : %8lld clk for %d copies -> %.1f B/clk/SM, %.0f clk per copy\n", shared_src, grid, bytes,
               depth, mx, iters, (double)iters * bytes / mx, (double)mx / iters);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
