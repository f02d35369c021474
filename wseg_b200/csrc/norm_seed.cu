// K1b — full-resolution min/max normalisation + bilinear(align_corners) resize + label mask, fwd and bwd
//        (replaces tool/visualization.py:62-67 and contrast_train.py:145-152 / 157-158).
// K1c — seed labels on the loss grid (replaces contrast_train.py:186-197 / 215-226).
// All HBM-bound; planes whose image-level label is 0 are never read (their output is exactly 0).
#include <algorithm>
#include <cfloat>
#include <climits>

#include "common.cuh"
#include "resample.cuh"

namespace wsegk {

struct MinMax {
    float lo, hi;
    int ilo, ihi;
};

__device__ __forceinline__ MinMax mm_merge(MinMax a, MinMax b) {
    // first flat index wins ties (torch.max/min(dim) backward routes to one index)
    MinMax r = a;
    if (b.lo < r.lo || (b.lo == r.lo && b.ilo < r.ilo)) { r.lo = b.lo; r.ilo = b.ilo; }
    if (b.hi > r.hi || (b.hi == r.hi && b.ihi < r.ihi)) { r.hi = b.hi; r.ihi = b.ihi; }
    return r;
}

__device__ __forceinline__ MinMax mm_warp(MinMax v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        MinMax w;
        w.lo = __shfl_xor_sync(0xffffffffu, v.lo, o);
        w.hi = __shfl_xor_sync(0xffffffffu, v.hi, o);
        w.ilo = __shfl_xor_sync(0xffffffffu, v.ilo, o);
        w.ihi = __shfl_xor_sync(0xffffffffu, v.ihi, o);
        v = mm_merge(v, w);
    }
    return v;
}

__device__ __forceinline__ MinMax mm_block(MinMax v, MinMax* s_red) {
    v = mm_warp(v);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    if (warp == 0) {
        MinMax w = lane < nw ? s_red[lane] : MinMax{FLT_MAX, -FLT_MAX, INT_MAX, INT_MAX};
        w = mm_warp(w);
        if (lane == 0) s_red[0] = w;
    }
    __syncthreads();
    MinMax r = s_red[0];
    __syncthreads();
    return r;
}

#ifndef WSEGK_PLANE_SPLIT
#define WSEGK_PLANE_SPLIT 8
#endif
#ifndef WSEGK_BWD_NBLK
#define WSEGK_BWD_NBLK 64
#endif
constexpr int kPlaneSplit = WSEGK_PLANE_SPLIT;   // CTAs per (n,c) plane in the min/max pass
constexpr int kUpPlaneSplit = 32;    // ... when the plane is evaluated on the fly (compute- instead of bandwidth-bound)
constexpr int kPlaneSplitMax = 32;   // what the workspace is sized for

// partial min/max (+ first arg) of relu(p) over a slice of one plane; masked planes exit immediately
__global__ void __launch_bounds__(512) plane_minmax_partial_kernel(const float* __restrict__ p,
                                                                   const float* __restrict__ label,
                                                                   MinMax* __restrict__ partial, int HW) {
    const int plane = blockIdx.y, part = blockIdx.x;
    __shared__ MinMax s_red[16];
    MinMax v{FLT_MAX, -FLT_MAX, INT_MAX, INT_MAX};
    if (label == nullptr || label[plane] != 0.f) {
        const float* src = p + static_cast<size_t>(plane) * HW;
        const int per = ceil_div_dev(HW, static_cast<int>(gridDim.x));
        const int per4 = (per + 3) & ~3;
        const int begin = part * per4, end = min(HW, begin + per4);
        const bool vec = ((HW & 3) == 0);
        if (vec) {
            const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
            for (int i = (begin >> 2) + threadIdx.x; i < (end >> 2); i += blockDim.x) {
                const float4 x = __ldg(s4 + i);
                const float e[4] = {fmaxf(x.x, 0.f), fmaxf(x.y, 0.f), fmaxf(x.z, 0.f), fmaxf(x.w, 0.f)};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int idx = 4 * i + j;
                    if (e[j] < v.lo) { v.lo = e[j]; v.ilo = idx; }
                    if (e[j] > v.hi) { v.hi = e[j]; v.ihi = idx; }
                }
            }
        } else {
            for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
                const float e = fmaxf(__ldg(src + i), 0.f);
                if (e < v.lo) { v.lo = e; v.ilo = i; }
                if (e > v.hi) { v.hi = e; v.ihi = i; }
            }
        }
    }
    v = mm_block(v, s_red);
    if (threadIdx.x == 0) partial[plane * gridDim.x + part] = v;
}

__device__ __forceinline__ MinMax plane_minmax_final(const MinMax* partial, int plane, int nsplit) {
    MinMax v = partial[plane * nsplit];
    for (int s = 1; s < nsplit; ++s) v = mm_merge(v, partial[plane * nsplit + s]);
    return v;
}

__device__ __forceinline__ float norm_val(float p, float lo, float inv_den) {
    return fmaxf(fmaxf(p, 0.f) - lo - 1e-5f, 0.f) * inv_den;
}

// out[n,c,oy,ox] = label * bilinear( relu(relu(p)-lo-e) / (hi-lo+e) )
__global__ void __launch_bounds__(256) fullres_resample_kernel(const float* __restrict__ p,
                                                               const float* __restrict__ label,
                                                               const MinMax* __restrict__ partial,
                                                               float* __restrict__ out,
                                                               float* __restrict__ minmax,
                                                               float* __restrict__ tapsum, int H, int W,
                                                               int oh, int ow, int nsplit) {
    const int plane = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const float lab = label[plane];
    float* dst = out + static_cast<size_t>(plane) * oh * ow;
    if (lab == 0.f) {
        if (o < oh * ow) {
            dst[o] = 0.f;
            if (tapsum) tapsum[static_cast<size_t>(plane) * oh * ow + o] = 0.f;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            minmax[plane * 4 + 0] = 0.f; minmax[plane * 4 + 1] = 0.f;
            minmax[plane * 4 + 2] = __int_as_float(0); minmax[plane * 4 + 3] = __int_as_float(0);
        }
        return;
    }
    const MinMax mm = plane_minmax_final(partial, plane, nsplit);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        minmax[plane * 4 + 0] = mm.lo; minmax[plane * 4 + 1] = mm.hi;
        minmax[plane * 4 + 2] = __int_as_float(mm.ilo); minmax[plane * 4 + 3] = __int_as_float(mm.ihi);
    }
    if (o >= oh * ow) return;
    const float den = mm.hi - mm.lo + 1e-5f;
    const float* src = p + static_cast<size_t>(plane) * H * W;
    float val, taps = 0.f;
    if (oh == H && ow == W) {
        val = fmaxf(fmaxf(__ldg(src + o), 0.f) - mm.lo - 1e-5f, 0.f) / den;
        taps = val > 0.f ? 1.f : 0.f;
    } else {
        const int oy = o / ow, ox = o - oy * ow;
        const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
        const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
        int y0, y1, x0, x1; float hy0, hy1, wx0, wx1;
        src_coord(oy, sy, H, y0, y1, hy0, hy1);
        src_coord(ox, sx, W, x0, x1, wx0, wx1);
        auto at = [&](int y, int x) {
            return fmaxf(fmaxf(__ldg(src + static_cast<size_t>(y) * W + x), 0.f) - mm.lo - 1e-5f, 0.f) / den;
        };
        const float v00 = at(y0, x0), v01 = at(y0, x1), v10 = at(y1, x0), v11 = at(y1, x1);
        val = bilerp(hy0, hy1, wx0, wx1, v00, v01, v10, v11);
        // sum of the tap weights whose (normalised) input is positive: what the backward needs to form
        // S0 = sum_pixels gy*[z>0]/den from the OUTPUT-sized tensors alone
        taps = bilerp(hy0, hy1, wx0, wx1, v00 > 0.f ? 1.f : 0.f, v01 > 0.f ? 1.f : 0.f, v10 > 0.f ? 1.f : 0.f,
                      v11 > 0.f ? 1.f : 0.f);
    }
    dst[o] = val * lab;
    if (tapsum) tapsum[static_cast<size_t>(plane) * oh * ow + o] = taps;
}

// ---- backward ----------------------------------------------------------------------------------
// gy(y,x) = sum over outputs (oy,ox) whose taps include (y,x) of weight * grad_out * label.  Shrinking (scale >= 1) makes
// i0(o) strictly increasing, so an input coordinate is the FIRST tap of at most one output and the SECOND tap of at
// most one output: two (output index, weight) slots per coordinate, built in shared memory without atomics.
struct TapSlot {
    int o;
    float w;
};

__device__ __forceinline__ void build_tap_table(TapSlot* tab, int in_size, int out_size, float scale) {
    for (int i = threadIdx.x; i < 2 * in_size; i += blockDim.x) tab[i] = TapSlot{0, 0.f};
    __syncthreads();
    for (int o = threadIdx.x; o < out_size; o += blockDim.x) {
        int i0, i1; float l0, l1;
        src_coord(o, scale, in_size, i0, i1, l0, l1);
        if (i1 == i0) {
            tab[2 * i0] = TapSlot{o, l0 + l1};
        } else {
            tab[2 * i0] = TapSlot{o, l0};
            tab[2 * i1 + 1] = TapSlot{o, l1};
        }
    }
    __syncthreads();
}

// Plane sums the gradient of max/min needs, from the OUTPUT-sized tensors only (kSumSplit CTAs per labelled plane):
//   S1 = sum_pix gy*relu(z)/den^2 = sum_o gout_o * out_o / den          (out_o = lab * sum_taps w*relu(z)/den)
//   S0 = sum_pix gy*[z>0]/den     = lab/den * sum_o gout_o * tapsum_o   (tapsum_o = sum_taps w*[z>0])
// Partials are combined in a fixed order by the consumer => deterministic.
constexpr int kSumSplit = 8;
__global__ void __launch_bounds__(256) fullres_bwd_sums_kernel(const float* __restrict__ label,
                                                               const float* __restrict__ out,
                                                               const float* __restrict__ tapsum,
                                                               const float* __restrict__ gout, float* __restrict__ sums,
                                                               int osz) {
    const int plane = blockIdx.y, part = blockIdx.x;
    if (label[plane] == 0.f) return;
    const float* o = out + static_cast<size_t>(plane) * osz;
    const float* t = tapsum + static_cast<size_t>(plane) * osz;
    const float* g = gout + static_cast<size_t>(plane) * osz;
    const int per = ceil_div_dev(osz, kSumSplit);
    const int begin = part * per, end = min(osz, begin + per);
    float a0 = 0.f, a1 = 0.f;
    int i = begin + threadIdx.x;
    for (; i + 3 * 256 < end; i += 4 * 256) {          // four independent triples of loads in flight
        float gv[4], tv[4], ov[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { gv[j] = __ldg(g + i + j * 256); tv[j] = __ldg(t + i + j * 256); ov[j] = __ldg(o + i + j * 256); }
#pragma unroll
        for (int j = 0; j < 4; ++j) { a0 = fmaf(gv[j], tv[j], a0); a1 = fmaf(gv[j], ov[j], a1); }
    }
    for (; i < end; i += 256) {
        const float gv = __ldg(g + i);
        a0 = fmaf(gv, __ldg(t + i), a0);
        a1 = fmaf(gv, __ldg(o + i), a1);
    }
    __shared__ float s_a[8], s_b[8];
#pragma unroll
    for (int ofs = 16; ofs > 0; ofs >>= 1) {
        a0 += __shfl_xor_sync(0xffffffffu, a0, ofs);
        a1 += __shfl_xor_sync(0xffffffffu, a1, ofs);
    }
    if ((threadIdx.x & 31) == 0) { s_a[threadIdx.x >> 5] = a0; s_b[threadIdx.x >> 5] = a1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float x0 = 0.f, x1 = 0.f;
        for (int wv = 0; wv < (blockDim.x >> 5); ++wv) { x0 += s_a[wv]; x1 += s_b[wv]; }
        sums[(plane * kSumSplit + part) * 2] = x0;
        sums[(plane * kSumSplit + part) * 2 + 1] = x1;
    }
}

// grad_p = [p>0] * ( gy*[z>0]/den + [idx==argmax]*(-S1) + [idx==argmin]*(S1 - S0) ); planes whose label is 0 are a
// vectorised zero fill.
__global__ void __launch_bounds__(256) fullres_bwd_kernel(const float* __restrict__ p,
                                                          const float* __restrict__ label,
                                                          const float* __restrict__ minmax,
                                                          const float* __restrict__ gout,
                                                          float* __restrict__ grad_p,
                                                          const float* __restrict__ sums, int H, int W, int oh,
                                                          int ow) {
    extern __shared__ TapSlot s_tab[];            // [2*H] rows then [2*W] columns (resampling case only)
    const int plane = blockIdx.y;
    const int HW = H * W;
    const float lab = label[plane];
    float* dst = grad_p + static_cast<size_t>(plane) * HW;
    if (lab == 0.f) {
        if ((HW & 3) == 0) {
            float4* d4 = reinterpret_cast<float4*>(dst);
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (HW >> 2); i += gridDim.x * blockDim.x)
                d4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        } else {
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) dst[i] = 0.f;
        }
        return;
    }
    const float lo = minmax[plane * 4], hi = minmax[plane * 4 + 1];
    const int ilo = __float_as_int(minmax[plane * 4 + 2]), ihi = __float_as_int(minmax[plane * 4 + 3]);
    const float inv_den = 1.f / (hi - lo + 1e-5f);
    const float* src = p + static_cast<size_t>(plane) * HW;
    const float* g = gout + static_cast<size_t>(plane) * oh * ow;
    const bool same = (oh == H && ow == W);
    TapSlot* ty = s_tab;
    TapSlot* tx = s_tab + 2 * H;
    if (!same) {
        const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
        const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
        build_tap_table(ty, H, oh, sy);
        build_tap_table(tx, W, ow, sx);
    }
    float S0 = 0.f, S1 = 0.f;
    for (int k = 0; k < kSumSplit; ++k) { S0 += sums[(plane * kSumSplit + k) * 2]; S1 += sums[(plane * kSumSplit + k) * 2 + 1]; }
    S0 *= lab * inv_den;
    S1 *= inv_den;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const float pv = __ldg(src + i);
        const float z = fmaxf(pv, 0.f) - lo - 1e-5f;
        float gy = 0.f;
        if (z > 0.f) {
            if (same) {
                gy = __ldg(g + i);
            } else {
                const int y = i / W, x = i - y * W;
                const TapSlot ya = ty[2 * y], yb = ty[2 * y + 1], xa = tx[2 * x], xb = tx[2 * x + 1];
                if (ya.w != 0.f) {
                    if (xa.w != 0.f) gy = fmaf(ya.w * xa.w, __ldg(g + ya.o * ow + xa.o), gy);
                    if (xb.w != 0.f) gy = fmaf(ya.w * xb.w, __ldg(g + ya.o * ow + xb.o), gy);
                }
                if (yb.w != 0.f) {
                    if (xa.w != 0.f) gy = fmaf(yb.w * xa.w, __ldg(g + yb.o * ow + xa.o), gy);
                    if (xb.w != 0.f) gy = fmaf(yb.w * xb.w, __ldg(g + yb.o * ow + xb.o), gy);
                }
            }
        }
        float d = gy * lab * inv_den;
        if (i == ihi) d -= S1;
        if (i == ilo) d += S1 - S0;
        dst[i] = (pv > 0.f) ? d : 0.f;
    }
}

// ---- glue: bilinear(align_corners) shrink of small maps (contrast_train.py:179-180) -----------------------------------
// ATen's NCHW bilinear kernel parallelises over the OUTPUT PIXELS only and loops over N*C inside a thread: for
// [8,128,56,56] -> 16x16 that is 256 threads and 0.3 ms.  One thread per output element here.
// enlargement with ow % 4 == 0: four consecutive outputs of a row per thread, one 16-byte store (the x8 case writes
// 135 MB per call and is bound by that store stream; the 12.5 KB source plane sits in L1)
__global__ void __launch_bounds__(256) resize_fwd_vec4_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              int planes, int H, int W, int oh, int ow) {
    const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
    const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
    const int ow4 = ow >> 2, osz4 = oh * ow4;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const float* src = x + static_cast<size_t>(plane) * H * W;
        float4* dst = reinterpret_cast<float4*>(y + static_cast<size_t>(plane) * oh * ow);
        for (int o4 = blockIdx.x * blockDim.x + threadIdx.x; o4 < osz4; o4 += gridDim.x * blockDim.x) {
            const int oy = o4 / ow4, ox0 = (o4 - oy * ow4) << 2;
            int y0, y1; float hy0, hy1;
            src_coord(oy, sy, H, y0, y1, hy0, hy1);
            const float* r0 = src + y0 * W;
            const float* r1 = src + y1 * W;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int x0, x1; float wx0, wx1;
                src_coord(ox0 + j, sx, W, x0, x1, wx0, wx1);
                v[j] = bilerp(hy0, hy1, wx0, wx1, __ldg(r0 + x0), __ldg(r0 + x1), __ldg(r1 + x0), __ldg(r1 + x1));
            }
            dst[o4] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
}

// The same for oh % RB == 0, a thread = four consecutive outputs of RB consecutive rows: the column taps and weights are
// computed once per thread and the row taps once per row, so a value costs ~12 instructions instead of ~40 and the kernel is
// bound by its store stream (the x8 enlargements at the end of Net.forward, network/resnet38_contrast.py:57-59, write
// 135 MB each at N = 8, 448x448: 85 -> ~25 us).  Same src_coord / bilerp arithmetic: bit-identical values.
template <int RB>
__global__ void __launch_bounds__(256) resize_fwd_rows_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                              int planes, int H, int W, int oh, int ow) {
    const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
    const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
    const int ow4 = ow >> 2, items = (oh / RB) * ow4;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const float* src = x + static_cast<size_t>(plane) * H * W;
        float4* dst = reinterpret_cast<float4*>(y + static_cast<size_t>(plane) * oh * ow);
        for (int it = blockIdx.x * blockDim.x + threadIdx.x; it < items; it += gridDim.x * blockDim.x) {
            const int rg = it / ow4, q = it - rg * ow4;
            int x0[4], x1[4];
            float wx0[4], wx1[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) src_coord(4 * q + j, sx, W, x0[j], x1[j], wx0[j], wx1[j]);
#pragma unroll
            for (int r = 0; r < RB; ++r) {
                const int oy = rg * RB + r;
                int y0, y1; float hy0, hy1;
                src_coord(oy, sy, H, y0, y1, hy0, hy1);
                const float* r0 = src + y0 * W;
                const float* r1 = src + y1 * W;
                float v[4];
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    v[j] = bilerp(hy0, hy1, wx0[j], wx1[j], __ldg(r0 + x0[j]), __ldg(r0 + x1[j]), __ldg(r1 + x0[j]), __ldg(r1 + x1[j]));
                dst[oy * ow4 + q] = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

__global__ void __launch_bounds__(256) resize_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int planes,
                                                         int H, int W, int oh, int ow) {
    const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
    const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
    const int osz = oh * ow;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const float* src = x + static_cast<size_t>(plane) * H * W;
        float* dst = y + static_cast<size_t>(plane) * osz;
        for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < osz; o += gridDim.x * blockDim.x) {
            const int oy = o / ow, ox = o - oy * ow;
            int y0, y1, x0, x1; float hy0, hy1, wx0, wx1;
            src_coord(oy, sy, H, y0, y1, hy0, hy1);
            src_coord(ox, sx, W, x0, x1, wx0, wx1);
            dst[o] = bilerp(hy0, hy1, wx0, wx1, __ldg(src + y0 * W + x0), __ldg(src + y0 * W + x1),
                            __ldg(src + y1 * W + x0), __ldg(src + y1 * W + x1));
        }
    }
}

// gx[plane,y,x] = sum of the (at most four) output gradients whose taps include (y,x); one CTA per plane group
__global__ void __launch_bounds__(256) resize_bwd_kernel(const float* __restrict__ gy, float* __restrict__ gx, int planes,
                                                         int H, int W, int oh, int ow) {
    extern __shared__ TapSlot s_tab[];
    TapSlot* ty = s_tab;
    TapSlot* tx = s_tab + 2 * H;
    const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
    const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
    build_tap_table(ty, H, oh, sy);
    build_tap_table(tx, W, ow, sx);
    const int isz = H * W;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const float* g = gy + static_cast<size_t>(plane) * oh * ow;
        float* dst = gx + static_cast<size_t>(plane) * isz;
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < isz; i += gridDim.x * blockDim.x) {
            const int yq = i / W, xq = i - yq * W;
            const TapSlot ya = ty[2 * yq], yb = ty[2 * yq + 1], xa = tx[2 * xq], xb = tx[2 * xq + 1];
            float acc = 0.f;
            if (ya.w != 0.f) {
                if (xa.w != 0.f) acc = fmaf(ya.w * xa.w, __ldg(g + ya.o * ow + xa.o), acc);
                if (xb.w != 0.f) acc = fmaf(ya.w * xb.w, __ldg(g + ya.o * ow + xb.o), acc);
            }
            if (yb.w != 0.f) {
                if (xa.w != 0.f) acc = fmaf(yb.w * xa.w, __ldg(g + yb.o * ow + xa.o), acc);
                if (xb.w != 0.f) acc = fmaf(yb.w * xb.w, __ldg(g + yb.o * ow + xb.o), acc);
            }
            dst[i] = acc;
        }
    }
}

constexpr int kMaxTapRange = 24;      // outputs touching one input coordinate along one axis (x8 enlargement: 18)
__global__ void __launch_bounds__(256) resize_bwd_gather_kernel(const float* __restrict__ gy, float* __restrict__ gx,
                                                                int planes, int H, int W, int oh, int ow,
                                                                const float* __restrict__ label, int c0_stride) {
    const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
    const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
    const int isz = H * W;
    for (int plane = blockIdx.y; plane < planes; plane += gridDim.y) {
        const float* g = gy + static_cast<size_t>(plane) * oh * ow;
        float* dst = gx + static_cast<size_t>(plane) * isz;
        if (label != nullptr && (label[plane] == 0.f || (c0_stride > 0 && plane % c0_stride == 0))) {
            // an unlabelled plane: its full-resolution gradient was never written and is not read
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < isz; i += gridDim.x * blockDim.x) dst[i] = 0.f;
            continue;
        }
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < isz; i += gridDim.x * blockDim.x) {
            const int yq = i / W, xq = i - yq * W;
            int oy0, oy1, ox0, ox1;
            out_range(yq, sy, oh, oy0, oy1);
            out_range(xq, sx, ow, ox0, ox1);
            float acc = 0.f;
            if (ox1 - ox0 < kMaxTapRange) {
                float wx[kMaxTapRange];                 // column weights once per input pixel
#pragma unroll
                for (int j = 0; j < kMaxTapRange; ++j) wx[j] = (ox0 + j <= ox1) ? tap_weight(ox0 + j, xq, sx, W) : 0.f;
                for (int oy = oy0; oy <= oy1; ++oy) {
                    const float wy = tap_weight(oy, yq, sy, H);
                    if (wy == 0.f) continue;
                    const float* row_p = g + static_cast<size_t>(oy) * ow + ox0;
                    float row = 0.f;
#pragma unroll
                    for (int j = 0; j < kMaxTapRange; ++j)
                        if (wx[j] != 0.f) row = fmaf(wx[j], __ldg(row_p + j), row);
                    acc = fmaf(wy, row, acc);
                }
            } else {
                for (int oy = oy0; oy <= oy1; ++oy) {
                    const float wy = tap_weight(oy, yq, sy, H);
                    if (wy == 0.f) continue;
                    float row = 0.f;
                    for (int ox = ox0; ox <= ox1; ++ox) {
                        const float wxv = tap_weight(ox, xq, sx, W);
                        if (wxv != 0.f) row = fmaf(wxv, __ldg(g + static_cast<size_t>(oy) * ow + ox), row);
                    }
                    acc = fmaf(wy, row, acc);
                }
            }
            dst[i] = acc;
        }
    }
}

// ---- K1c ---------------------------------------------------------------------------------------
// one warp per (n,c) plane: min/max of relu(x)
__global__ void __launch_bounds__(256) seed_minmax_kernel(const float* __restrict__ x, float* __restrict__ lohi,
                                                          int planes, int HW) {
    const int plane = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (plane >= planes) return;
    const int lane = threadIdx.x & 31;
    const float* src = x + static_cast<size_t>(plane) * HW;
    float lo = FLT_MAX, hi = 0.f;
    for (int i = lane; i < HW; i += 32) {
        const float v = fmaxf(__ldg(src + i), 0.f);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if (lane == 0) { lohi[2 * plane] = lo; lohi[2 * plane + 1] = hi; }
}

// thread per pixel: normalise every class plane, write norm_cam, argmax_c(norm*label) with lowest class on ties
__global__ void __launch_bounds__(128) seed_pixel_kernel(const float* __restrict__ x, const float* __restrict__ label,
                                                         const float* __restrict__ lohi, float bg,
                                                         float* __restrict__ norm_cam, int32_t* __restrict__ pseudo,
                                                         int C, int HW) {
    const int n = blockIdx.y;
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ float s_lo[32], s_hi[32], s_lab[32];
    if (threadIdx.x < C) {
        s_lo[threadIdx.x] = lohi[2 * (n * C + threadIdx.x)];
        s_hi[threadIdx.x] = lohi[2 * (n * C + threadIdx.x) + 1];
        s_lab[threadIdx.x] = label[n * C + threadIdx.x];
    }
    __syncthreads();
    if (q >= HW) return;
    const float* src = x + static_cast<size_t>(n) * C * HW + q;
    float* dst = norm_cam + static_cast<size_t>(n) * C * HW + q;
    float best = bg * s_lab[0];
    int arg = 0;
    dst[0] = bg;
    for (int c0 = 1; c0 < C; c0 += 8) {                       // eight independent loads in flight
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = (c0 + j < C) ? __ldg(src + static_cast<size_t>(c0 + j) * HW) : 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            if (c < C) {
                float v = fmaxf(x[j], 0.f);
                const float lo = s_lo[c], hi = s_hi[c];
                if (v < lo + 1e-5f) v = 0.f;
                v = (v - lo - 1e-5f) / (hi - lo + 1e-5f);
                dst[static_cast<size_t>(c) * HW] = v;
                const float s = v * s_lab[c];
                if (s > best) { best = s; arg = c; }
            }
        }
    }
    pseudo[n * HW + q] = arg;
}

// ---- K1b on a LAZY full-resolution map ------------------------------------------------------------
// Net.forward enlarges the stride-8 maps x8 (network/resnet38_contrast.py:57-59) and the loss then reads the 135 MB result
// back to normalise and shrink it (contrast_train.py:145-152).  These kernels take the SMALL map and evaluate the enlarged
// one on the fly — the same expression, in the same order, as resize_fwd_kernel, so every value is bit-identical to the
// materialised tensor — and the backward gathers straight into the small map's gradient.  The full-resolution tensor, its
// dense gradient and the zero planes of both never exist.
__global__ void __launch_bounds__(512) up_minmax_partial_kernel(const float* __restrict__ small,
                                                                const float* __restrict__ label,
                                                                MinMax* __restrict__ partial, UpGeom g) {
    extern __shared__ float s_plane[];
    const int plane = blockIdx.y, part = blockIdx.x;
    __shared__ MinMax s_red[16];
    MinMax v{FLT_MAX, -FLT_MAX, INT_MAX, INT_MAX};
    if (label[plane] != 0.f) {                                 // uniform per CTA
        load_small_plane(small + static_cast<size_t>(plane) * g.h * g.w, s_plane, g.h * g.w);
        const int HW = g.H * g.W;
        const int per = ceil_div_dev(HW, static_cast<int>(gridDim.x));
        const int begin = part * per, end = min(HW, begin + per);
        for (int i = begin + threadIdx.x; i < end; i += blockDim.x) {
            const int y = i / g.W, x = i - y * g.W;
            const float e = fmaxf(up_value(s_plane, g, y, x), 0.f);
            if (e < v.lo) { v.lo = e; v.ilo = i; }
            if (e > v.hi) { v.hi = e; v.ihi = i; }
        }
    }
    v = mm_block(v, s_red);
    if (threadIdx.x == 0) partial[plane * gridDim.x + part] = v;
}

__global__ void __launch_bounds__(256) up_resample_kernel(const float* __restrict__ small,
                                                          const float* __restrict__ label,
                                                          const MinMax* __restrict__ partial, float* __restrict__ out,
                                                          float* __restrict__ minmax, float* __restrict__ tapsum,
                                                          UpGeom g, int oh, int ow, int nsplit) {
    extern __shared__ float s_plane[];
    const int plane = blockIdx.y;
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const float lab = label[plane];
    float* dst = out + static_cast<size_t>(plane) * oh * ow;
    if (lab == 0.f) {
        if (o < oh * ow) {
            dst[o] = 0.f;
            tapsum[static_cast<size_t>(plane) * oh * ow + o] = 0.f;
        }
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            minmax[plane * 4 + 0] = 0.f; minmax[plane * 4 + 1] = 0.f;
            minmax[plane * 4 + 2] = __int_as_float(0); minmax[plane * 4 + 3] = __int_as_float(0);
        }
        return;
    }
    load_small_plane(small + static_cast<size_t>(plane) * g.h * g.w, s_plane, g.h * g.w);
    const MinMax mm = plane_minmax_final(partial, plane, nsplit);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        minmax[plane * 4 + 0] = mm.lo; minmax[plane * 4 + 1] = mm.hi;
        minmax[plane * 4 + 2] = __int_as_float(mm.ilo); minmax[plane * 4 + 3] = __int_as_float(mm.ihi);
    }
    if (o >= oh * ow) return;
    const float den = mm.hi - mm.lo + 1e-5f;
    const int H = g.H, W = g.W;
    auto at = [&](int y, int x) { return fmaxf(fmaxf(up_value(s_plane, g, y, x), 0.f) - mm.lo - 1e-5f, 0.f) / den; };
    float val, taps;
    if (oh == H && ow == W) {
        const int y = o / W, x = o - y * W;
        val = at(y, x);
        taps = val > 0.f ? 1.f : 0.f;
    } else {
        const int oy = o / ow, ox = o - oy * ow;
        const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
        const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
        int y0, y1, x0, x1; float hy0, hy1, wx0, wx1;
        src_coord(oy, sy, H, y0, y1, hy0, hy1);
        src_coord(ox, sx, W, x0, x1, wx0, wx1);
        const float v00 = at(y0, x0), v01 = at(y0, x1), v10 = at(y1, x0), v11 = at(y1, x1);
        val = bilerp(hy0, hy1, wx0, wx1, v00, v01, v10, v11);
        taps = bilerp(hy0, hy1, wx0, wx1, v00 > 0.f ? 1.f : 0.f, v01 > 0.f ? 1.f : 0.f, v10 > 0.f ? 1.f : 0.f,
                      v11 > 0.f ? 1.f : 0.f);
    }
    dst[o] = val * lab;
    tapsum[static_cast<size_t>(plane) * oh * ow + o] = taps;
}

// Backward: the full-resolution gradient d(i) = [p_i>0] * ( gy_i*[z_i>0]*lab/den + [i==argmax]*(-S1) + [i==argmin]*(S1-S0) )
// (as fullres_bwd_kernel, with p_i re-evaluated from the small plane in shared memory) is written to a scratch tensor for
// the LABELLED planes only — about 2.5 of 21 per image; the other planes of the scratch are never touched — and the
// enlargement's transpose (resize_bwd_gather_kernel, which skips the same planes) brings it onto the small map.
__global__ void __launch_bounds__(256) up_dense_grad_kernel(const float* __restrict__ small, const float* __restrict__ label,
                                                            const float* __restrict__ minmax,
                                                            const float* __restrict__ gout, float* __restrict__ scratch,
                                                            const float* __restrict__ sums, UpGeom g, int oh, int ow) {
    extern __shared__ float s_dyn_up[];
    const int plane = blockIdx.y;
    const float lab = label[plane];
    if (lab == 0.f) return;
    const int hw = g.h * g.w, H = g.H, W = g.W, HW = H * W;
    float* s_plane = s_dyn_up;
    TapSlot* ty = reinterpret_cast<TapSlot*>(s_dyn_up + ((hw + 1) & ~1));
    TapSlot* tx = ty + 2 * H;
    load_small_plane(small + static_cast<size_t>(plane) * hw, s_plane, hw);
    const bool same = (oh == H && ow == W);
    if (!same) {
        const float sy = oh > 1 ? static_cast<float>(H - 1) / static_cast<float>(oh - 1) : 0.f;
        const float sx = ow > 1 ? static_cast<float>(W - 1) / static_cast<float>(ow - 1) : 0.f;
        build_tap_table(ty, H, oh, sy);
        build_tap_table(tx, W, ow, sx);
    }
    const float lo = minmax[plane * 4], hi = minmax[plane * 4 + 1];
    const int ilo = __float_as_int(minmax[plane * 4 + 2]), ihi = __float_as_int(minmax[plane * 4 + 3]);
    const float inv_den = 1.f / (hi - lo + 1e-5f);
    float S0 = 0.f, S1 = 0.f;
    for (int k = 0; k < kSumSplit; ++k) { S0 += sums[(plane * kSumSplit + k) * 2]; S1 += sums[(plane * kSumSplit + k) * 2 + 1]; }
    S0 *= lab * inv_den;
    S1 *= inv_den;
    const float* gp = gout + static_cast<size_t>(plane) * oh * ow;
    float* dst = scratch + static_cast<size_t>(plane) * HW;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const int y = i / W, x = i - y * W;
        const float pv = up_value(s_plane, g, y, x);
        const float z = fmaxf(pv, 0.f) - lo - 1e-5f;
        float gy = 0.f;
        if (z > 0.f) {
            if (same) {
                gy = __ldg(gp + i);
            } else {
                const TapSlot ya = ty[2 * y], yb = ty[2 * y + 1], xa = tx[2 * x], xb = tx[2 * x + 1];
                if (ya.w != 0.f) {
                    if (xa.w != 0.f) gy = fmaf(ya.w * xa.w, __ldg(gp + ya.o * ow + xa.o), gy);
                    if (xb.w != 0.f) gy = fmaf(ya.w * xb.w, __ldg(gp + ya.o * ow + xb.o), gy);
                }
                if (yb.w != 0.f) {
                    if (xa.w != 0.f) gy = fmaf(yb.w * xa.w, __ldg(gp + yb.o * ow + xa.o), gy);
                    if (xb.w != 0.f) gy = fmaf(yb.w * xb.w, __ldg(gp + yb.o * ow + xb.o), gy);
                }
            }
        }
        float d = gy * lab * inv_den;
        if (i == ihi) d -= S1;
        if (i == ilo) d += S1 - S0;
        dst[i] = (pv > 0.f) ? d : 0.f;
    }
}

int launch_resize_bwd_labelled(const float* gy, const float* label, float* gx, int planes, int H, int W, int oh, int ow,
                               int c0_stride, cudaStream_t stream) {
    const dim3 grid(std::min(ceil_div(H * W, 256), 1024), std::min(planes, 32768));
    resize_bwd_gather_kernel<<<grid, 256, 0, stream>>>(gy, gx, planes, H, W, oh, ow, label, c0_stride);
    count_launches(1);
    return WSEGK_OK;
}

}  // namespace wsegk

using namespace wsegk;

static int fullres_check(const char* who, int N, int C, int H, int W, int oh, int ow) {
    WSEGK_REQUIRE(N > 0 && C > 0 && H > 0 && W > 0 && oh > 0 && ow > 0, WSEGK_ERR_BAD_SHAPE,
                  "%s: all dims must be positive", who);
    WSEGK_REQUIRE(oh <= H && ow <= W, WSEGK_ERR_BAD_SHAPE,
                  "%s: output (%d,%d) larger than input (%d,%d) is not supported (the reference only shrinks)", who,
                  oh, ow, H, W);
    WSEGK_REQUIRE(static_cast<long long>(H) * W < (1ll << 30), WSEGK_ERR_BAD_SHAPE, "%s: plane too large", who);
    return WSEGK_OK;
}

static size_t fullres_ws_bytes(int N, int C) {
    // min/max partials + backward block sums
    return align_up(static_cast<size_t>(N) * C * kPlaneSplitMax * sizeof(MinMax), 1024) +
           align_up(static_cast<size_t>(N) * C * 64 * 2 * sizeof(float), 1024);
}

namespace wsegk {
size_t fullres_workspace_bytes(int N, int C) { return fullres_ws_bytes(N, C); }
size_t lazy_bwd_workspace_bytes(int N, int C, int H, int W) {
    return fullres_ws_bytes(N, C) + align_up(static_cast<size_t>(N) * C * H * W * sizeof(float), 1024);
}
}

extern "C" int wsegk_fullres_norm_fwd(const float* p, const float* label, float* out, float* minmax, float* tapsum,
                                      int N, int C, int H, int W, int oh, int ow, void* workspace,
                                      size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(p && label && out && minmax && workspace, WSEGK_ERR_NULL, "fullres_norm_fwd: null pointer");
    if (int rc = fullres_check("fullres_norm_fwd", N, C, H, W, oh, ow)) return rc;
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(p) & 15) == 0, WSEGK_ERR_BAD_ALIGN, "fullres_norm_fwd: p unaligned");
    WSEGK_REQUIRE(workspace_bytes >= fullres_ws_bytes(N, C), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "fullres_norm_fwd: workspace %zu < %zu", workspace_bytes, fullres_ws_bytes(N, C));
    MinMax* partial = static_cast<MinMax*>(workspace);
    const int planes = N * C;
    plane_minmax_partial_kernel<<<dim3(kPlaneSplit, planes), kPlaneSplit > 8 ? 256 : 512, 0, stream>>>(p, label, partial, H * W);
    fullres_resample_kernel<<<dim3(ceil_div(oh * ow, 256), planes), 256, 0, stream>>>(p, label, partial, out, minmax,
                                                                                       tapsum, H, W, oh, ow, kPlaneSplit);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_fullres_norm_bwd(const float* p, const float* label, const float* minmax, const float* out,
                                      const float* tapsum, const float* grad_out, float* grad_p, int N, int C, int H,
                                      int W, int oh, int ow, void* workspace, size_t workspace_bytes,
                                      wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(p && label && minmax && out && tapsum && grad_out && grad_p && workspace, WSEGK_ERR_NULL,
                  "fullres_norm_bwd: null pointer");
    if (int rc = fullres_check("fullres_norm_bwd", N, C, H, W, oh, ow)) return rc;
    WSEGK_REQUIRE(workspace_bytes >= fullres_ws_bytes(N, C), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "fullres_norm_bwd: workspace %zu < %zu", workspace_bytes, fullres_ws_bytes(N, C));
    float* sums = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) +
                                           align_up(static_cast<size_t>(N) * C * kPlaneSplitMax * sizeof(MinMax), 1024));
    const int planes = N * C;
    const int nblk = std::min(WSEGK_BWD_NBLK, ceil_div(H * W, 256 * 4));
    const size_t tab = (oh == H && ow == W) ? 0 : static_cast<size_t>(2) * (H + W) * sizeof(TapSlot);
    WSEGK_REQUIRE(tab <= 200 * 1024, WSEGK_ERR_BAD_SHAPE, "fullres_norm_bwd: H + W = %d too large for the tap tables",
                  H + W);
    WSEGK_CUDA_OK(cudaFuncSetAttribute(fullres_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(tab)));
    fullres_bwd_sums_kernel<<<dim3(kSumSplit, planes), 256, 0, stream>>>(label, out, tapsum, grad_out, sums, oh * ow);
    fullres_bwd_kernel<<<dim3(nblk, planes), 256, tab, stream>>>(p, label, minmax, grad_out, grad_p, sums, H, W, oh, ow);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

static int up_check(const char* who, int N, int C, int h, int w, int H, int W, int oh, int ow) {
    if (int rc = fullres_check(who, N, C, H, W, oh, ow)) return rc;
    WSEGK_REQUIRE(h > 0 && w > 0 && h <= H && w <= W, WSEGK_ERR_BAD_SHAPE,
                  "%s: the small map (%dx%d) must not be larger than the full-resolution one (%dx%d)", who, h, w, H, W);
    WSEGK_REQUIRE(static_cast<size_t>(h) * w * 4 + static_cast<size_t>(2) * (H + W) * sizeof(TapSlot) + 64 <= 200 * 1024,
                  WSEGK_ERR_BAD_SHAPE, "%s: small plane + tap tables do not fit shared memory", who);
    return WSEGK_OK;
}
extern "C" int wsegk_fullres_norm_up_fwd(const float* small, const float* label, float* out, float* minmax, float* tapsum,
                                         int N, int C, int h, int w, int H, int W, int oh, int ow, void* workspace,
                                         size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(small && label && out && minmax && tapsum && workspace, WSEGK_ERR_NULL, "fullres_norm_up_fwd: null pointer");
    if (int rc = up_check("fullres_norm_up_fwd", N, C, h, w, H, W, oh, ow)) return rc;
    WSEGK_REQUIRE(workspace_bytes >= fullres_ws_bytes(N, C), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "fullres_norm_up_fwd: workspace %zu < %zu", workspace_bytes, fullres_ws_bytes(N, C));
    MinMax* partial = static_cast<MinMax*>(workspace);
    const int planes = N * C;
    const UpGeom g = up_geom(h, w, H, W);
    const size_t smem = static_cast<size_t>(h) * w * sizeof(float);
    WSEGK_CUDA_OK(cudaFuncSetAttribute(up_minmax_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    WSEGK_CUDA_OK(cudaFuncSetAttribute(up_resample_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    up_minmax_partial_kernel<<<dim3(kUpPlaneSplit, planes), 256, smem, stream>>>(small, label, partial, g);
    up_resample_kernel<<<dim3(ceil_div(oh * ow, 256), planes), 256, smem, stream>>>(small, label, partial, out, minmax, tapsum,
                                                                                  g, oh, ow, kUpPlaneSplit);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_fullres_norm_up_bwd(const float* small, const float* label, const float* minmax, const float* out,
                                         const float* tapsum, const float* grad_out, float* grad_small, int N, int C,
                                         int h, int w, int H, int W, int oh, int ow, void* workspace,
                                         size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(small && label && minmax && out && tapsum && grad_out && grad_small && workspace, WSEGK_ERR_NULL,
                  "fullres_norm_up_bwd: null pointer");
    if (int rc = up_check("fullres_norm_up_bwd", N, C, h, w, H, W, oh, ow)) return rc;
    WSEGK_REQUIRE(workspace_bytes >= lazy_bwd_workspace_bytes(N, C, H, W), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "fullres_norm_up_bwd: workspace %zu < %zu (WSEGK_OP_LAZY_BWD)", workspace_bytes,
                  lazy_bwd_workspace_bytes(N, C, H, W));
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    float* sums = reinterpret_cast<float*>(ws + align_up(static_cast<size_t>(N) * C * kPlaneSplitMax * sizeof(MinMax), 1024));
    float* scratch = reinterpret_cast<float*>(ws + fullres_ws_bytes(N, C));
    const int planes = N * C;
    const UpGeom g = up_geom(h, w, H, W);
    const bool same = (oh == H && ow == W);
    const size_t smem = static_cast<size_t>((h * w + 1) & ~1) * sizeof(float) +
                        (same ? 0 : static_cast<size_t>(2) * (H + W) * sizeof(TapSlot));
    WSEGK_CUDA_OK(cudaFuncSetAttribute(up_dense_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    fullres_bwd_sums_kernel<<<dim3(kSumSplit, planes), 256, 0, stream>>>(label, out, tapsum, grad_out, sums, oh * ow);
    const int nblk = std::min(64, ceil_div(H * W, 256 * 4));
    up_dense_grad_kernel<<<dim3(nblk, planes), 256, smem, stream>>>(small, label, minmax, grad_out, scratch, sums, g, oh, ow);
    count_launches(2);
    launch_resize_bwd_labelled(scratch, label, grad_small, planes, h, w, H, W, 0, stream);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_seed_labels_fwd(const float* cam_rv_down, const float* label, float bg_threshold,
                                     float* norm_cam, int32_t* pseudo, int N, int C, int h, int w, void* workspace,
                                     size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(cam_rv_down && label && norm_cam && pseudo && workspace, WSEGK_ERR_NULL,
                  "seed_labels_fwd: null pointer");
    WSEGK_REQUIRE(N > 0 && h > 0 && w > 0 && C >= 1 && C <= 32, WSEGK_ERR_BAD_SHAPE,
                  "seed_labels_fwd: need N,h,w > 0 and 1 <= C <= 32");
    const size_t need = wsegk_workspace_bytes(WSEGK_OP_SEED_LABELS, N, C, 0, 0, h, w);
    WSEGK_REQUIRE(workspace_bytes >= need, WSEGK_ERR_WORKSPACE_TOO_SMALL, "seed_labels_fwd: workspace %zu < %zu",
                  workspace_bytes, need);
    float* lohi = static_cast<float*>(workspace);
    const int planes = N * C, HW = h * w;
    seed_minmax_kernel<<<ceil_div(planes, 8), 256, 0, stream>>>(cam_rv_down, lohi, planes, HW);
    seed_pixel_kernel<<<dim3(ceil_div(HW, 128), N), 128, 0, stream>>>(cam_rv_down, label, lohi, bg_threshold,
                                                                      norm_cam, pseudo, C, HW);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

static int resize_check(const char* who, int planes, int H, int W, int oh, int ow) {
    WSEGK_REQUIRE(planes > 0 && H > 0 && W > 0 && oh > 0 && ow > 0, WSEGK_ERR_BAD_SHAPE, "%s: dims must be positive", who);
    WSEGK_REQUIRE(static_cast<long long>(H) * W < (1ll << 30) && static_cast<long long>(oh) * ow < (1ll << 30),
                  WSEGK_ERR_BAD_SHAPE, "%s: plane too large", who);
    return WSEGK_OK;
}

extern "C" int wsegk_bilinear_resize_fwd(const float* x, float* y, int planes, int H, int W, int oh, int ow,
                                         wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(x && y, WSEGK_ERR_NULL, "bilinear_resize_fwd: null pointer");
    if (int rc = resize_check("bilinear_resize_fwd", planes, H, W, oh, ow)) return rc;
    if ((ow & 3) == 0 && (oh & 7) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && oh * ow >= 4096 && oh >= H && ow >= W) {
        const dim3 grid(std::min(ceil_div((oh >> 3) * (ow >> 2), 256), 1024), std::min(planes, 32768));
        resize_fwd_rows_kernel<8><<<grid, 256, 0, stream>>>(x, y, planes, H, W, oh, ow);
    } else if ((ow & 3) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 && oh * ow >= 4096) {
        const dim3 grid(std::min(ceil_div(oh * (ow >> 2), 256), 1024), std::min(planes, 32768));
        resize_fwd_vec4_kernel<<<grid, 256, 0, stream>>>(x, y, planes, H, W, oh, ow);
    } else {
        const dim3 grid(std::min(ceil_div(oh * ow, 256), 1024), std::min(planes, 32768));
        resize_fwd_kernel<<<grid, 256, 0, stream>>>(x, y, planes, H, W, oh, ow);
    }
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_bilinear_resize_bwd(const float* grad_y, float* grad_x, int planes, int H, int W, int oh, int ow,
                                         wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(grad_y && grad_x, WSEGK_ERR_NULL, "bilinear_resize_bwd: null pointer");
    if (int rc = resize_check("bilinear_resize_bwd", planes, H, W, oh, ow)) return rc;
    const dim3 grid(std::min(ceil_div(H * W, 256), 1024), std::min(planes, 32768));
    const size_t tab = static_cast<size_t>(2) * (H + W) * sizeof(TapSlot);
    if (oh <= H && ow <= W && tab <= 200 * 1024) {      // shrinking: <= 4 contributors per input, tap tables
        WSEGK_CUDA_OK(cudaFuncSetAttribute(resize_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(tab)));
        resize_bwd_kernel<<<grid, 256, tab, stream>>>(grad_y, grad_x, planes, H, W, oh, ow);
    } else {
        resize_bwd_gather_kernel<<<grid, 256, 0, stream>>>(grad_y, grad_x, planes, H, W, oh, ow, nullptr, 0);
    }
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}
