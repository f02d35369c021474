// Bilinear (align_corners=True) resampling helpers shared by the materialising resize kernels, the lazy K1b kernels
// (norm_seed.cu) and the lazy classification / min-pool kernels (aux_loss.cu).  One definition of every rounding step, so a
// value computed on the fly is bit-identical to the one a materialised tensor would hold.
#pragma once
#include "common.cuh"

namespace wsegk {

// bilinear, align_corners=True source coordinate (ATen area_pixel_compute_source_index)
__device__ __forceinline__ void src_coord(int o, float scale, int in_size, int& i0, int& i1, float& l0, float& l1) {
    const float s = scale * static_cast<float>(o);
    i0 = static_cast<int>(s);
    if (i0 > in_size - 1) i0 = in_size - 1;
    i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
    l1 = s - static_cast<float>(i0);
    l0 = 1.f - l1;
}

// hy0*(wx0*a + wx1*b) + hy1*(wx0*c + wx1*d) with the contraction into FMAs spelled out, so that every kernel that
// evaluates a bilinear sample (the materialising resize, the lazy K1b kernels) rounds identically.
__device__ __forceinline__ float bilerp(float hy0, float hy1, float wx0, float wx1, float a, float b, float c, float d) {
    const float top = __fmaf_rn(wx0, a, __fmul_rn(wx1, b));
    const float bot = __fmaf_rn(wx0, c, __fmul_rn(wx1, d));
    return __fmaf_rn(hy0, top, __fmul_rn(hy1, bot));
}

// general gather (any scale, e.g. the x8 enlargement of Net.forward, network/resnet38_contrast.py:57-59): an input
// coordinate i is touched by the outputs whose first tap is i-1 or i; first taps are monotone in o, so they form the
// contiguous range below (padded by one on each side against float rounding); weights come from the forward's own
// arithmetic (tap_weight), so the result is the exact transpose of the forward.
__device__ __forceinline__ float tap_weight(int o, int i, float scale, int in_size) {
    int i0, i1; float l0, l1;
    src_coord(o, scale, in_size, i0, i1, l0, l1);
    float wgt = 0.f;
    if (i0 == i) wgt += l0;
    if (i1 == i) wgt += l1;
    return wgt;
}

__device__ __forceinline__ void out_range(int i, float scale, int out_size, int& first, int& last) {
    if (scale <= 0.f) { first = 0; last = out_size - 1; return; }
    const int a = static_cast<int>(floorf(static_cast<float>(i - 1) / scale)) - 1;
    const int b = static_cast<int>(ceilf(static_cast<float>(i + 1) / scale)) + 1;
    first = max(a, 0);
    last = min(b, out_size - 1);
}

// ---- a full-resolution map evaluated on the fly from the small (stride-8) one --------------------------------------
struct UpGeom {
    int h, w, H, W;
    float sy, sx;              // (h-1)/(H-1), (w-1)/(W-1): source step of the enlargement
};
__device__ __forceinline__ float up_value(const float* __restrict__ sm, const UpGeom& g, int y, int x) {
    int y0, y1, x0, x1; float hy0, hy1, wx0, wx1;
    src_coord(y, g.sy, g.h, y0, y1, hy0, hy1);
    src_coord(x, g.sx, g.w, x0, x1, wx0, wx1);
    return bilerp(hy0, hy1, wx0, wx1, sm[y0 * g.w + x0], sm[y0 * g.w + x1], sm[y1 * g.w + x0], sm[y1 * g.w + x1]);
}
// n floats global -> shared, eight independent loads per thread in flight (a rolled load->store loop costs one L2 latency
// per iteration); ends with a __syncthreads()
__device__ __forceinline__ void stage_plane(const float* __restrict__ src, float* dst, int n) {
    const int step = static_cast<int>(blockDim.x);
    for (int i0 = threadIdx.x; i0 < n; i0 += 8 * step) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (i0 + j * step < n) ? __ldg(src + i0 + j * step) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (i0 + j * step < n) dst[i0 + j * step] = v[j];
    }
    __syncthreads();
}
__device__ __forceinline__ void load_small_plane(const float* __restrict__ src, float* s_plane, int hw) {
    stage_plane(src, s_plane, hw);
}

inline UpGeom up_geom(int h, int w, int H, int W) {
    UpGeom g;
    g.h = h; g.w = w; g.H = H; g.W = W;
    g.sy = H > 1 ? static_cast<float>(h - 1) / static_cast<float>(H - 1) : 0.f;      // as resize_fwd_kernel
    g.sx = W > 1 ? static_cast<float>(w - 1) / static_cast<float>(W - 1) : 0.f;
    return g;
}


}  // namespace wsegk
