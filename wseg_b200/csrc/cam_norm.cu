// K1a — in-network CAM normalisation (replaces network/resnet38_contrast.py:41-48).
//   x = relu(cam); peak[n,c] = max_hw x + 1e-5; y = relu(x - 1e-5) / peak
//   y[:,0] = 1 - max_{c>=1} y;  for c>=1: y[c] = (y[c] < max_{c>=1} y) ? 0 : y[c]
// HBM-bound: one read for the plane maxima, one read + one write for the per-pixel pass.
#include "common.cuh"

namespace wsegk {

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one CTA per (n,c) plane
__global__ void __launch_bounds__(256) cam_plane_peak_kernel(const float* __restrict__ cam,
                                                             float* __restrict__ peak, int HW) {
    const float* src = cam + static_cast<size_t>(blockIdx.x) * HW;
    float m = 0.f;   // relu => max >= 0
    if ((HW & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
        for (int i = threadIdx.x; i < (HW >> 2); i += blockDim.x) {
            float4 v = __ldg(s4 + i);
            m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
        }
    } else {
        for (int i = threadIdx.x; i < HW; i += blockDim.x) m = fmaxf(m, __ldg(src + i));
    }
    __shared__ float red[8];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        v = warp_max(v);
        if (threadIdx.x == 0) peak[blockIdx.x] = v + 1e-5f;
    }
}

// one thread per pixel; channels are strided by HW so every load/store is coalesced across the warp.  Two short
// passes over the channels (the second re-reads from L1/L2) keep the code small: at these sizes (2 MB) the kernel is
// bound by launch and instruction-fetch latency, and a 32-way unrolled body took 13 us.
template <int MAXC>
__global__ void __launch_bounds__(128) cam_pixel_kernel(const float* __restrict__ cam,
                                                        const float* __restrict__ peak,
                                                        float* __restrict__ out, int C, int HW) {
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float s_inv[MAXC];
    if (threadIdx.x < C) s_inv[threadIdx.x] = peak[n * C + threadIdx.x];
    __syncthreads();
    if (p >= HW) return;
    const float* src = cam + static_cast<size_t>(n) * C * HW + p;
    float* dst = out + static_cast<size_t>(n) * C * HW + p;
    float fg = 0.f;
#pragma unroll 4
    for (int c = 1; c < C; ++c) {
        const float x = fmaxf(__ldg(src + static_cast<size_t>(c) * HW), 0.f);
        fg = fmaxf(fg, fmaxf(x - 1e-5f, 0.f) / s_inv[c]);
    }
    dst[0] = 1.f - fg;
#pragma unroll 4
    for (int c = 1; c < C; ++c) {
        const float x = fmaxf(__ldg(src + static_cast<size_t>(c) * HW), 0.f);
        const float y = fmaxf(x - 1e-5f, 0.f) / s_inv[c];
        dst[static_cast<size_t>(c) * HW] = (y < fg) ? 0.f : y;
    }
}

void launch_cam_plane_peak(const float* cam, float* peak, int planes, int HW, cudaStream_t stream) {
    cam_plane_peak_kernel<<<planes, 256, 0, stream>>>(cam, peak, HW);
}

}  // namespace wsegk

using namespace wsegk;

extern "C" int wsegk_cam_norm_fwd(const float* cam, float* cam_norm, int N, int C, int h, int w,
                                  void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(cam && cam_norm && workspace, WSEGK_ERR_NULL, "cam_norm_fwd: null pointer");
    WSEGK_REQUIRE(N > 0 && h > 0 && w > 0 && C >= 2 && C <= 32, WSEGK_ERR_BAD_SHAPE,
                  "cam_norm_fwd: need N,h,w > 0 and 2 <= C <= 32 (got N=%d C=%d h=%d w=%d)", N, C, h, w);
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(cam) & 15) == 0, WSEGK_ERR_BAD_ALIGN,
                  "cam_norm_fwd: cam must be 16-byte aligned");
    const size_t need = wsegk_workspace_bytes(WSEGK_OP_CAM_NORM, N, C, 0, 0, h, w);
    WSEGK_REQUIRE(workspace_bytes >= need, WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "cam_norm_fwd: workspace %zu < %zu", workspace_bytes, need);
    const int HW = h * w;
    float* peak = static_cast<float*>(workspace);
    cam_plane_peak_kernel<<<N * C, 256, 0, stream>>>(cam, peak, HW);
    dim3 grid(ceil_div(HW, 128), N);
    cam_pixel_kernel<32><<<grid, 128, 0, stream>>>(cam, peak, cam_norm, C, HW);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}
