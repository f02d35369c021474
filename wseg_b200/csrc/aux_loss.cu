// Auxiliary SEAM losses of the contrast step (the first "next" row):
//   classification  : GAP + multilabel soft margin                  contrast_train.py:142,155,159-160
//   min-pool        : adaptive_min_pooling_loss (bottom 25 % of the per-pixel class maximum)   :16-25,143,156
//   ER              : mean |cam1 - cam2| over the foreground classes                           :163
//   ECR             : mean of the top 20 % of |max_onehot(cam_other) - cam_rv| per image        :28-32,165-172
// The two "top-k of a few hundred thousand values per image" are done by value: a radix select finds the k-th value,
// the loss is a sum over "beyond the threshold" plus the tied share, and the gradient is an indicator — no sort, no
// index tensor.  Every floating-point reduction runs in a fixed order (no float atomics).
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "resample.cuh"

namespace wsegk {

__device__ __forceinline__ uint32_t aux_ordered_bits(float v) {
    if (v == 0.f) v = 0.f;                       // -0 -> +0
    const uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float aux_unordered_bits(uint32_t o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
// larger key == selected earlier
__device__ __forceinline__ uint32_t sel_key(float v, bool largest) {
    const uint32_t o = aux_ordered_bits(v);
    return largest ? o : ~o;
}

// fixed-order block sum (blockDim.x multiple of 32, <= 1024); every thread gets the result
__device__ __forceinline__ float block_sum(float v, float* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += s_red[w];
    return t;
}
__device__ __forceinline__ int block_sum_int(int v, int* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    int t = 0;
    for (int w = 0; w < nw; ++w) t += s_red[w];
    return t;
}

// ------------------------------------------------------------------------------------------------
// select: the k-th value of each row, by value
// ------------------------------------------------------------------------------------------------
// out[row] = { threshold t, sum of the k selected values, sum of max(selected, 0), (float) number of ties taken };
// idx_cut[row]: elements equal to t are selected iff their index is < idx_cut (== len when all ties are taken).
// Element i with value v is selected iff  key(v) > key(t)  ||  (v == t && i < idx_cut[row]).
//
// A row (200 704 or 344 064 values) is split over S CTAs; a digit pass is two small launches: every CTA histograms its
// chunk in shared memory and adds the non-empty bins to the row's global histogram (integer atomics: deterministic),
// then one warp per row picks the digit.  (A single CTA per row took 218 us per call.)
constexpr int kSelThreads = 256;
constexpr int kSelBins = 2048;
constexpr int kSelMaxSplit = 64;

struct SelState {
    uint32_t prefix;
    int need;
};

__global__ void sel_init_kernel(SelState* __restrict__ st, uint32_t* __restrict__ hist, int rows, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows) { st[i].prefix = 0u; st[i].need = k; }
    for (int j = i; j < rows * kSelBins; j += gridDim.x * blockDim.x) hist[j] = 0u;
}

__device__ __forceinline__ void sel_chunk(int len, int nsplit, int part, int& begin, int& end) {
    const int per = ((ceil_div_dev(len, nsplit) + 3) & ~3);
    begin = min(len, part * per);
    end = min(len, begin + per);
}

__global__ void __launch_bounds__(kSelThreads) sel_hist_kernel(const float* __restrict__ x, int len, int largest, int shift,
                                                               int width, int first, const SelState* __restrict__ st,
                                                               uint32_t* __restrict__ hist) {
    __shared__ uint32_t s_hist[kSelBins];
    const int row = blockIdx.y;
    const float* src = x + static_cast<size_t>(row) * len;
    const bool lg = largest != 0;
    int begin, end;
    sel_chunk(len, gridDim.x, blockIdx.x, begin, end);
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x) s_hist[b] = 0u;
    __syncthreads();
    const uint32_t prefix = st[row].prefix;
    for (int base = begin; base < end; base += 8 * kSelThreads) {          // warp-uniform trip count (ballots inside)
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = base + threadIdx.x + j * kSelThreads;
            v[j] = i < end ? __ldg(src + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool in = base + threadIdx.x + j * kSelThreads < end;
            const uint32_t key = sel_key(v[j], lg);
            const bool act = in && (first || (key >> (shift + width)) == prefix);
            const uint32_t bin = (key >> shift) & ((1u << width) - 1u);
            if (first) {
                // every element takes part and the leading digit (sign, exponent, 2 mantissa bits) takes few values:
                // aggregate equal bins inside the warp before touching shared memory
                const unsigned am = __ballot_sync(0xffffffffu, act);
                if (act) {
                    const unsigned peers = __match_any_sync(am, bin);
                    if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], static_cast<uint32_t>(__popc(peers)));
                }
            } else if (act) {
                atomicAdd(&s_hist[bin], 1u);               // few survivors, digits spread out
            }
        }
    }
    __syncthreads();
    uint32_t* g = hist + static_cast<size_t>(row) * kSelBins;
    for (int b = threadIdx.x; b < kSelBins; b += blockDim.x)
        if (s_hist[b]) atomicAdd(g + b, s_hist[b]);
}

// one CTA per row: the bin (descending) holding the need-th key; clears the histogram for the next pass.  Thread t owns
// the 8 bins nb-1-8t .. nb-8-8t (one warp walking all 2048 bins cost 7.6 us per pass and row set).
__global__ void __launch_bounds__(256) sel_scan_kernel(SelState* __restrict__ st, uint32_t* __restrict__ hist, int width) {
    __shared__ int s_warp[8];
    const int row = blockIdx.x, t = threadIdx.x, lane = t & 31, warp = t >> 5;
    uint32_t* g = hist + static_cast<size_t>(row) * kSelBins;
    const int nb = 1 << width;
    int local[8], sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int bin = nb - 1 - (8 * t + j);
        local[j] = bin >= 0 ? static_cast<int>(g[bin]) : 0;
        sum += local[j];
    }
    const int need = st[row].need;
    const uint32_t prefix = st[row].prefix;
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int excl = incl - sum;
    for (int w = 0; w < warp; ++w) excl += s_warp[w];
    if (excl < need && need <= excl + sum) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (cum + local[j] >= need) {
                st[row].need = need - cum;
                st[row].prefix = (prefix << width) | static_cast<uint32_t>(nb - 1 - (8 * t + j));
                break;
            }
            cum += local[j];
        }
    }
    for (int b = t; b < kSelBins; b += 256) g[b] = 0u;
}

// per chunk: sum / relu-sum over the strictly-beyond elements and the number of ties
__global__ void __launch_bounds__(kSelThreads) sel_sum_kernel(const float* __restrict__ x, int len, int largest,
                                                              const SelState* __restrict__ st, float4* __restrict__ partial) {
    __shared__ float s_redf[32];
    __shared__ int s_redi[32];
    const int row = blockIdx.y;
    const float* src = x + static_cast<size_t>(row) * len;
    const bool lg = largest != 0;
    int begin, end;
    sel_chunk(len, gridDim.x, blockIdx.x, begin, end);
    const uint32_t tkey = st[row].prefix;
    float sum = 0.f, sum_relu = 0.f;
    int ceq = 0;
    for (int base = begin; base < end; base += 8 * kSelThreads) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int i = base + threadIdx.x + j * kSelThreads;
            v[j] = i < end ? __ldg(src + i) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (base + threadIdx.x + j * kSelThreads < end) {
                const uint32_t key = sel_key(v[j], lg);
                if (key > tkey) { sum += v[j]; sum_relu += fmaxf(v[j], 0.f); }
                ceq += key == tkey ? 1 : 0;
            }
        }
    }
    sum = block_sum(sum, s_redf);
    sum_relu = block_sum(sum_relu, s_redf);
    ceq = block_sum_int(ceq, s_redi);
    if (threadIdx.x == 0)
        partial[row * kSelMaxSplit + blockIdx.x] = make_float4(sum, sum_relu, __int_as_float(ceq), 0.f);
}

// one CTA per row: fixed-order sum of the chunk partials; if there are more ties than needed, an ordered scan of the row
// finds the index cut (lowest indices first)
__global__ void __launch_bounds__(1024) sel_final_kernel(const float* __restrict__ x, int len, int largest, int nsplit,
                                                         const SelState* __restrict__ st,
                                                         const float4* __restrict__ partial, float4* __restrict__ out,
                                                         int* __restrict__ idx_cut) {
    __shared__ int s_redi[32];
    __shared__ int s_cut;
    const int row = blockIdx.x;
    const float* src = x + static_cast<size_t>(row) * len;
    const bool lg = largest != 0;
    const uint32_t tkey = st[row].prefix;
    const int need = st[row].need;
    const float t = aux_unordered_bits(lg ? tkey : ~tkey);
    float sum = 0.f, sum_relu = 0.f;
    int ceq = 0;
    for (int p = 0; p < nsplit; ++p) {                         // every thread: same order, same result
        const float4 q = partial[row * kSelMaxSplit + p];
        sum += q.x; sum_relu += q.y; ceq += __float_as_int(q.z);
    }
    if (threadIdx.x == 0) s_cut = len;
    __syncthreads();
    if (ceq > need) {
        // the chunk partials already count the ties of every (contiguous) chunk: skip to the chunk holding the need-th tie
        int seen = 0, begin = 0, end = len;
        for (int p = 0; p < nsplit; ++p) {
            const int ties = __float_as_int(partial[row * kSelMaxSplit + p].z);
            sel_chunk(len, nsplit, p, begin, end);
            if (seen + ties >= need) break;
            seen += ties;
        }
        for (int base = begin; base < end; base += 1024) {
            const int i = base + threadIdx.x;
            const int flag = (i < end && sel_key(__ldg(src + i), lg) == tkey) ? 1 : 0;
            int incl = flag;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, incl, o);
                if ((threadIdx.x & 31) >= o) incl += u;
            }
            __syncthreads();
            if ((threadIdx.x & 31) == 31) s_redi[threadIdx.x >> 5] = incl;
            __syncthreads();
            int woff = 0, total = 0;
            for (int w = 0; w < 32; ++w) {
                if (w < (threadIdx.x >> 5)) woff += s_redi[w];
                total += s_redi[w];
            }
            const int rank = seen + woff + incl;               // 1-based rank of this tie among all ties
            if (flag && rank == need) s_cut = i + 1;
            seen += total;
            __syncthreads();
            if (seen >= need) break;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        out[row] = make_float4(t, sum + static_cast<float>(need) * t, sum_relu + static_cast<float>(need) * fmaxf(t, 0.f),
                               static_cast<float>(need));
        idx_cut[row] = s_cut;
    }
}

static size_t select_ws_bytes(int rows) {
    return align_up(static_cast<size_t>(rows) * kSelBins * sizeof(uint32_t), 256) +
           align_up(static_cast<size_t>(rows) * sizeof(SelState), 256) +
           align_up(static_cast<size_t>(rows) * kSelMaxSplit * sizeof(float4), 256);
}
size_t select_workspace_bytes(int rows) { return select_ws_bytes(rows); }

static int select_launch(const float* x, int rows, int len, int k, int largest, float* sel, int32_t* idx_cut, void* workspace,
                         wsegk_stream_t stream) {
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    uint32_t* hist = reinterpret_cast<uint32_t*>(ws);
    SelState* st = reinterpret_cast<SelState*>(ws + align_up(static_cast<size_t>(rows) * kSelBins * sizeof(uint32_t), 256));
    float4* partial = reinterpret_cast<float4*>(reinterpret_cast<uint8_t*>(st) +
                                                align_up(static_cast<size_t>(rows) * sizeof(SelState), 256));
    const int nsplit = std::max(1, std::min(kSelMaxSplit, ceil_div(len, 8192)));
    sel_init_kernel<<<std::min(64, ceil_div(rows * kSelBins, 256)), 256, 0, stream>>>(st, hist, rows, k);
    const int shifts[3] = {21, 10, 0}, widths[3] = {11, 11, 10};
    for (int pass = 0; pass < 3; ++pass) {
        sel_hist_kernel<<<dim3(nsplit, rows), kSelThreads, 0, stream>>>(x, len, largest, shifts[pass], widths[pass],
                                                                        pass == 0 ? 1 : 0, st, hist);
        sel_scan_kernel<<<rows, 256, 0, stream>>>(st, hist, widths[pass]);
    }
    sel_sum_kernel<<<dim3(nsplit, rows), kSelThreads, 0, stream>>>(x, len, largest, st, partial);
    sel_final_kernel<<<rows, 1024, 0, stream>>>(x, len, largest, nsplit, st, partial, reinterpret_cast<float4*>(sel), idx_cut);
    count_launches(9);
    return WSEGK_OK;
}

__device__ __forceinline__ bool is_selected(float v, int i, float t, int cut, bool largest) {
    const uint32_t kv = sel_key(v, largest), kt = sel_key(t, largest);
    return kv > kt || (kv == kt && i < cut);
}

// ------------------------------------------------------------------------------------------------
// classification: GAP + multilabel soft margin
// ------------------------------------------------------------------------------------------------
constexpr int kGapSplit = 8;
__global__ void __launch_bounds__(512) gap_partial_kernel(const float* __restrict__ cam, float* __restrict__ partial, int HW) {
    __shared__ float s_red[32];
    const int plane = blockIdx.y, part = blockIdx.x;
    const float* src = cam + static_cast<size_t>(plane) * HW;
    const int per = ((ceil_div_dev(HW, kGapSplit) + 3) & ~3);
    const int begin = part * per, end = min(HW, begin + per);
    float acc = 0.f;
    if ((HW & 3) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
        for (int i = (begin >> 2) + threadIdx.x; i < (end >> 2); i += blockDim.x) {
            const float4 v = __ldg(s4 + i);
            acc += (v.x + v.y) + (v.z + v.w);
        }
    } else {
        for (int i = begin + threadIdx.x; i < end; i += blockDim.x) acc += __ldg(src + i);
    }
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0) partial[plane * kGapSplit + part] = acc;
}

// one CTA: gap[n,c], loss = mean_n mean_{c>=1} -(y log sig(x) + (1-y) log sig(-x)), dgap = d loss / d gap
__global__ void __launch_bounds__(256) gap_mlsm_kernel(const float* __restrict__ partial, const float* __restrict__ label,
                                                       float* __restrict__ gap, float* __restrict__ dgap,
                                                       float* __restrict__ loss, int N, int C, int HW) {
    __shared__ float s_red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < N * C; i += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < kGapSplit; ++p) s += partial[i * kGapSplit + p];
        const float x = s / static_cast<float>(HW);
        gap[i] = x;
        const int c = i % C;
        float d = 0.f;
        if (c >= 1) {
            const float y = label[i];
            // log sigmoid(x) = min(x,0) - log1p(exp(-|x|))
            const float ls_pos = fminf(x, 0.f) - log1pf(expf(-fabsf(x)));
            const float ls_neg = fminf(-x, 0.f) - log1pf(expf(-fabsf(x)));
            acc += -(y * ls_pos + (1.f - y) * ls_neg);
            const float sig = 1.f / (1.f + expf(-x));
            d = (sig - y) / (static_cast<float>(C - 1) * static_cast<float>(N));
        }
        dgap[i] = d;
    }
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0) loss[0] = acc / (static_cast<float>(C - 1) * static_cast<float>(N));
}

// grad_cam[n,c,:] = dgap[n,c] * g / HW
__global__ void __launch_bounds__(256) gap_bwd_kernel(const float* __restrict__ dgap, const float* __restrict__ g,
                                                      float* __restrict__ grad_cam, int HW) {
    const int plane = blockIdx.y;
    const float v = dgap[plane] * g[0] / static_cast<float>(HW);
    float* dst = grad_cam + static_cast<size_t>(plane) * HW;
    if ((HW & 3) == 0) {
        float4* d4 = reinterpret_cast<float4*>(dst);
        const float4 v4 = make_float4(v, v, v, v);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < (HW >> 2); i += gridDim.x * blockDim.x) d4[i] = v4;
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) dst[i] = v;
    }
}

// ------------------------------------------------------------------------------------------------
// min-pool
// ------------------------------------------------------------------------------------------------
// m[n,p] = max_{c>=1} cam_rv[n,c,p]*label[n,c] (first index on ties, like torch.max(dim)); arg[n,p] = that class
__global__ void __launch_bounds__(256) minpool_max_kernel(const float* __restrict__ x, const float* __restrict__ label,
                                                          float* __restrict__ m, int32_t* __restrict__ arg, int C,
                                                          int HW) {
    __shared__ float s_lab[32];
    const int n = blockIdx.y;
    if (threadIdx.x < C) s_lab[threadIdx.x] = label[n * C + threadIdx.x];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const float* src = x + static_cast<size_t>(n) * C * HW + p;
    float best = -FLT_MAX;
    int bi = 1;
    for (int c0 = 1; c0 < C; c0 += 8) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            // planes with label 0 contribute exactly 0 and are not read
            v[j] = (c < C && s_lab[c] != 0.f) ? __ldg(src + static_cast<size_t>(c) * HW) * s_lab[c] : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = c0 + j;
            if (c < C && v[j] > best) { best = v[j]; bi = c; }
        }
    }
    m[static_cast<size_t>(n) * HW + p] = best;
    arg[static_cast<size_t>(n) * HW + p] = bi;
}

// grad_x[n,c,p] = [selected(p) && m>0 && c == arg] * label[n,c] * g / (k*N)      (dense write, channel 0 = 0)
__global__ void __launch_bounds__(256) minpool_bwd_kernel(const float* __restrict__ m, const int32_t* __restrict__ arg,
                                                          const float4* __restrict__ sel, const int* __restrict__ cut,
                                                          const float* __restrict__ label, const float* __restrict__ g,
                                                          float* __restrict__ grad_x, int N, int C, int HW, int k) {
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const float mv = m[static_cast<size_t>(n) * HW + p];
    const int a = arg[static_cast<size_t>(n) * HW + p];
    const bool on = mv > 0.f && is_selected(mv, p, sel[n].x, cut[n], false);
    const float val = on ? label[n * C + a] * g[0] / (static_cast<float>(k) * static_cast<float>(N)) : 0.f;
    float* dst = grad_x + static_cast<size_t>(n) * C * HW + p;
    for (int c = 0; c < C; ++c) dst[static_cast<size_t>(c) * HW] = (c == a) ? val : 0.f;
}

// ------------------------------------------------------------------------------------------------
// the same two losses on a LAZY full-resolution map: the small map goes in, the x8 enlargement of
// Net.forward (network/resnet38_contrast.py:57-59) is evaluated on the fly with the shared bilerp(), nothing of size
// N*C*H*W is read or written, and the gradients come back on the small map
// ------------------------------------------------------------------------------------------------
// column / row sums of the enlargement's weights: wsum[a] = sum over outputs o of tap_weight(o, a)
__device__ __forceinline__ float up_weight_sum(int a, float scale, int in_size, int out_size) {
    int first, last;
    out_range(a, scale, out_size, first, last);
    float acc = 0.f;
    for (int o = first; o <= last; ++o) acc += tap_weight(o, a, scale, in_size);
    return acc;
}

// GAP of the enlarged plane = sum_a wy[ya]*wx[xa]*small[a] / (H*W)  (the enlargement is linear).  One CTA per plane; the
// result goes to partial slot 0 of the plane so gap_mlsm_kernel is reused as is.
__global__ void __launch_bounds__(256) gap_up_kernel(const float* __restrict__ small, float* __restrict__ partial, UpGeom g) {
    extern __shared__ float s_w[];                     // wy[h], wx[w]
    __shared__ float s_red[32];
    const int plane = blockIdx.x;
    for (int i = threadIdx.x; i < g.h + g.w; i += blockDim.x)
        s_w[i] = i < g.h ? up_weight_sum(i, g.sy, g.h, g.H) : up_weight_sum(i - g.h, g.sx, g.w, g.W);
    __syncthreads();
    const float* src = small + static_cast<size_t>(plane) * g.h * g.w;
    float acc = 0.f;
    for (int a = threadIdx.x; a < g.h * g.w; a += blockDim.x) {
        const int ya = a / g.w, xa = a - ya * g.w;
        acc = fmaf(s_w[ya] * s_w[g.h + xa], __ldg(src + a), acc);
    }
    acc = block_sum(acc, s_red);
    if (threadIdx.x < kGapSplit) partial[plane * kGapSplit + threadIdx.x] = threadIdx.x == 0 ? acc : 0.f;
}

// grad_small[n,c,a] = dgap[n,c] * g / (H*W) * wy[ya]*wx[xa]
__global__ void __launch_bounds__(256) gap_up_bwd_kernel(const float* __restrict__ dgap, const float* __restrict__ gl,
                                                         float* __restrict__ grad_small, UpGeom g) {
    extern __shared__ float s_w[];
    const int plane = blockIdx.x;
    for (int i = threadIdx.x; i < g.h + g.w; i += blockDim.x)
        s_w[i] = i < g.h ? up_weight_sum(i, g.sy, g.h, g.H) : up_weight_sum(i - g.h, g.sx, g.w, g.W);
    __syncthreads();
    const float v = dgap[plane] * gl[0] / (static_cast<float>(g.H) * static_cast<float>(g.W));
    float* dst = grad_small + static_cast<size_t>(plane) * g.h * g.w;
    for (int a = threadIdx.x; a < g.h * g.w; a += blockDim.x) {
        const int ya = a / g.w, xa = a - ya * g.w;
        dst[a] = v * (s_w[ya] * s_w[g.h + xa]);
    }
}

// m[n,p] = max_{c>=1} up(small[n,c])[p]*label[n,c] over the full-resolution pixels p (first class on ties), arg = class
__global__ void __launch_bounds__(256) minpool_max_up_kernel(const float* __restrict__ small, const float* __restrict__ label,
                                                             float* __restrict__ m, int32_t* __restrict__ arg, int C,
                                                             UpGeom g) {
    __shared__ float s_lab[32];
    const int n = blockIdx.y;
    if (threadIdx.x < C) s_lab[threadIdx.x] = label[n * C + threadIdx.x];
    __syncthreads();
    const int HW = g.H * g.W, hw = g.h * g.w;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const int y = p / g.W, x = p - y * g.W;
    int y0, y1, x0, x1; float hy0, hy1, wx0, wx1;
    src_coord(y, g.sy, g.h, y0, y1, hy0, hy1);
    src_coord(x, g.sx, g.w, x0, x1, wx0, wx1);
    const int i00 = y0 * g.w + x0, i01 = y0 * g.w + x1, i10 = y1 * g.w + x0, i11 = y1 * g.w + x1;
    const float* src = small + static_cast<size_t>(n) * C * hw;
    float best = -FLT_MAX;
    int bi = 1;
    for (int c = 1; c < C; ++c) {
        float v = 0.f;                                   // planes with label 0 contribute exactly 0 and are not read
        if (s_lab[c] != 0.f) {
            const float* pl = src + static_cast<size_t>(c) * hw;
            v = bilerp(hy0, hy1, wx0, wx1, __ldg(pl + i00), __ldg(pl + i01), __ldg(pl + i10), __ldg(pl + i11)) * s_lab[c];
        }
        if (v > best) { best = v; bi = c; }
    }
    m[static_cast<size_t>(n) * HW + p] = best;
    arg[static_cast<size_t>(n) * HW + p] = bi;
}

// Backward: the full-resolution gradient [selected(p) && m>0 && c == arg] * label*g/(k*N) goes to a scratch tensor for the
// planes of the LABELLED foreground classes only (the rest of the scratch is never touched), then the enlargement's
// transpose — which skips the same planes and channel 0 — brings it onto the small map.
__global__ void __launch_bounds__(256) minpool_dense_labelled_kernel(const float* __restrict__ m, const int32_t* __restrict__ arg,
                                                                     const float4* __restrict__ sel, const int* __restrict__ cut,
                                                                     const float* __restrict__ label, const float* __restrict__ gl,
                                                                     float* __restrict__ scratch, int N, int C, int HW, int k) {
    __shared__ float s_lab[32];
    const int n = blockIdx.y;
    if (threadIdx.x < C) s_lab[threadIdx.x] = label[n * C + threadIdx.x];
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    const float mv = m[static_cast<size_t>(n) * HW + p];
    const int a = arg[static_cast<size_t>(n) * HW + p];
    const bool on = mv > 0.f && is_selected(mv, p, sel[n].x, cut[n], false);
    const float val = on ? s_lab[a] * gl[0] / (static_cast<float>(k) * static_cast<float>(N)) : 0.f;
    float* dst = scratch + static_cast<size_t>(n) * C * HW + p;
    for (int c = 1; c < C; ++c)
        if (s_lab[c] != 0.f) dst[static_cast<size_t>(c) * HW] = (c == a) ? val : 0.f;
}

// ------------------------------------------------------------------------------------------------
// ER + ECR on the 128x128 maps
// ------------------------------------------------------------------------------------------------
constexpr int kAuxMaxC = 24;
// e1 = |max_onehot(with_bg(n2)) - rv1|, e2 = |max_onehot(with_bg(n1)) - rv2|  ([N, C*hw] each);  er partial sums
__global__ void __launch_bounds__(128) er_ecr_fwd_kernel(const float* __restrict__ n1, const float* __restrict__ n2,
                                                         const float* __restrict__ rv1, const float* __restrict__ rv2,
                                                         float* __restrict__ e1, float* __restrict__ e2,
                                                         float* __restrict__ er_partial, int C, int hw) {
    __shared__ float s_red[32];
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    float er = 0.f;
    if (p < hw) {
        const size_t base = static_cast<size_t>(n) * C * hw + p;
        float a[kAuxMaxC], b[kAuxMaxC];
        float ma = -FLT_MAX, mb = -FLT_MAX;
#pragma unroll
        for (int c = 0; c < kAuxMaxC; ++c) {
            a[c] = 0.f; b[c] = 0.f;
            if (c < C) {
                a[c] = __ldg(n1 + base + static_cast<size_t>(c) * hw);
                b[c] = __ldg(n2 + base + static_cast<size_t>(c) * hw);
                if (c >= 1) { ma = fmaxf(ma, a[c]); mb = fmaxf(mb, b[c]); er += fabsf(a[c] - b[c]); }
            }
        }
#pragma unroll
        for (int c = 0; c < kAuxMaxC; ++c) {
            if (c < C) {
                const float oh1 = c == 0 ? 1.f - ma : (a[c] != ma ? 0.f : a[c]);
                const float oh2 = c == 0 ? 1.f - mb : (b[c] != mb ? 0.f : b[c]);
                e1[base + static_cast<size_t>(c) * hw] = fabsf(oh2 - __ldg(rv1 + base + static_cast<size_t>(c) * hw));
                e2[base + static_cast<size_t>(c) * hw] = fabsf(oh1 - __ldg(rv2 + base + static_cast<size_t>(c) * hw));
            }
        }
    }
    er = block_sum(er, s_red);
    if (threadIdx.x == 0) er_partial[blockIdx.y * gridDim.x + blockIdx.x] = er;
}

// out[0] = loss_er, out[1] = loss_ecr (fixed-order sums)
__global__ void __launch_bounds__(256) er_ecr_finalize_kernel(const float* __restrict__ er_partial, int nparts,
                                                              const float4* __restrict__ sel1, const float4* __restrict__ sel2,
                                                              float* __restrict__ out, int N, int C, int hw, int k) {
    __shared__ float s_red[32];
    float acc = 0.f;
    for (int i = threadIdx.x; i < nparts; i += blockDim.x) acc += er_partial[i];
    acc = block_sum(acc, s_red);
    if (threadIdx.x == 0) {
        out[0] = acc / (static_cast<float>(N) * static_cast<float>(C - 1) * static_cast<float>(hw));
        float s1 = 0.f, s2 = 0.f;
        for (int n = 0; n < N; ++n) { s1 += sel1[n].y; s2 += sel2[n].y; }
        out[1] = (s1 + s2) / (static_cast<float>(N) * static_cast<float>(k));
    }
}

// gradients of loss_er (w.r.t. n1, n2, channels >= 1) and loss_ecr (w.r.t. rv1, rv2)
__global__ void __launch_bounds__(128) er_ecr_bwd_kernel(const float* __restrict__ n1, const float* __restrict__ n2,
                                                         const float* __restrict__ rv1, const float* __restrict__ rv2,
                                                         const float4* __restrict__ sel1, const int* __restrict__ cut1,
                                                         const float4* __restrict__ sel2, const int* __restrict__ cut2,
                                                         const float* __restrict__ g_er, const float* __restrict__ g_ecr,
                                                         float* __restrict__ gn1, float* __restrict__ gn2,
                                                         float* __restrict__ grv1, float* __restrict__ grv2, int N, int C,
                                                         int hw, int k) {
    const int n = blockIdx.y;
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= hw) return;
    const size_t base = static_cast<size_t>(n) * C * hw + p;
    const float w_er = g_er[0] / (static_cast<float>(N) * static_cast<float>(C - 1) * static_cast<float>(hw));
    const float w_ecr = g_ecr[0] / (static_cast<float>(N) * static_cast<float>(k));
    const float t1 = sel1[n].x, t2 = sel2[n].x;
    const int c1 = cut1[n], c2 = cut2[n];
    float a[kAuxMaxC], b[kAuxMaxC];
    float ma = -FLT_MAX, mb = -FLT_MAX;
#pragma unroll
    for (int c = 0; c < kAuxMaxC; ++c) {
        a[c] = 0.f; b[c] = 0.f;
        if (c < C) {
            a[c] = __ldg(n1 + base + static_cast<size_t>(c) * hw);
            b[c] = __ldg(n2 + base + static_cast<size_t>(c) * hw);
            if (c >= 1) { ma = fmaxf(ma, a[c]); mb = fmaxf(mb, b[c]); }
        }
    }
#pragma unroll
    for (int c = 0; c < kAuxMaxC; ++c) {
        if (c < C) {
            const size_t off = base + static_cast<size_t>(c) * hw;
            const float d = a[c] - b[c];
            const float s = c >= 1 ? (d > 0.f ? w_er : (d < 0.f ? -w_er : 0.f)) : 0.f;
            gn1[off] = s;
            gn2[off] = -s;
            const float oh1 = c == 0 ? 1.f - ma : (a[c] != ma ? 0.f : a[c]);
            const float oh2 = c == 0 ? 1.f - mb : (b[c] != mb ? 0.f : b[c]);
            const float x1 = __ldg(rv1 + off), x2 = __ldg(rv2 + off);
            const float d1 = oh2 - x1, d2 = oh1 - x2;          // e = |d|, de/dx = -sign(d)
            const int idx = c * hw + p;                        // position inside the image's flattened [C*hw] row
            const bool on1 = is_selected(fabsf(d1), idx, t1, c1, true);
            const bool on2 = is_selected(fabsf(d2), idx, t2, c2, true);
            grv1[off] = on1 ? (d1 > 0.f ? -w_ecr : (d1 < 0.f ? w_ecr : 0.f)) : 0.f;
            grv2[off] = on2 ? (d2 > 0.f ? -w_ecr : (d2 < 0.f ? w_ecr : 0.f)) : 0.f;
        }
    }
}

}  // namespace wsegk

using namespace wsegk;

extern "C" int wsegk_select_kth(const float* x, int rows, int len, int k, int largest, float* sel, int32_t* idx_cut,
                                void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(x && sel && idx_cut && workspace, WSEGK_ERR_NULL, "select_kth: null pointer");
    WSEGK_REQUIRE(rows > 0 && len > 0 && k >= 1 && k <= len, WSEGK_ERR_BAD_SHAPE,
                  "select_kth: need rows, len > 0 and 1 <= k <= len (rows=%d len=%d k=%d)", rows, len, k);
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(sel) & 15) == 0 && (reinterpret_cast<uintptr_t>(workspace) & 255) == 0,
                  WSEGK_ERR_BAD_ALIGN, "select_kth: sel must be 16-byte and workspace 256-byte aligned");
    WSEGK_REQUIRE(workspace_bytes >= select_ws_bytes(rows), WSEGK_ERR_WORKSPACE_TOO_SMALL, "select_kth: workspace %zu < %zu",
                  workspace_bytes, select_ws_bytes(rows));
    select_launch(x, rows, len, k, largest, sel, idx_cut, workspace, stream);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_gap_mlsm_fwd(const float* cam, const float* label, float* gap, float* dgap, float* loss, int N, int C,
                                  int H, int W, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(cam && label && gap && dgap && loss && workspace, WSEGK_ERR_NULL, "gap_mlsm_fwd: null pointer");
    WSEGK_REQUIRE(N > 0 && C >= 2 && H > 0 && W > 0 && static_cast<long long>(H) * W < (1ll << 30), WSEGK_ERR_BAD_SHAPE,
                  "gap_mlsm_fwd: bad dims");
    WSEGK_REQUIRE(workspace_bytes >= static_cast<size_t>(N) * C * kGapSplit * sizeof(float), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "gap_mlsm_fwd: workspace too small");
    float* partial = static_cast<float*>(workspace);
    gap_partial_kernel<<<dim3(kGapSplit, N * C), 512, 0, stream>>>(cam, partial, H * W);
    gap_mlsm_kernel<<<1, 256, 0, stream>>>(partial, label, gap, dgap, loss, N, C, H * W);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_gap_mlsm_bwd(const float* dgap, const float* grad_loss, float* grad_cam, int N, int C, int H, int W,
                                  wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(dgap && grad_loss && grad_cam, WSEGK_ERR_NULL, "gap_mlsm_bwd: null pointer");
    WSEGK_REQUIRE(N > 0 && C >= 2 && H > 0 && W > 0, WSEGK_ERR_BAD_SHAPE, "gap_mlsm_bwd: bad dims");
    const int HW = H * W;
    gap_bwd_kernel<<<dim3(std::min(64, ceil_div(HW, 1024)), N * C), 256, 0, stream>>>(dgap, grad_loss, grad_cam, HW);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_minpool_fwd(const float* x, const float* label, float* m, int32_t* arg, float* sel, int32_t* idx_cut,
                                 int N, int C, int H, int W, void* workspace, size_t workspace_bytes,
                                 wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(x && label && m && arg && sel && idx_cut && workspace, WSEGK_ERR_NULL, "minpool_fwd: null pointer");
    WSEGK_REQUIRE(N > 0 && C >= 2 && C <= 32 && H > 0 && W > 0 && H * W >= 4, WSEGK_ERR_BAD_SHAPE, "minpool_fwd: bad dims");
    const int HW = H * W;
    minpool_max_kernel<<<dim3(ceil_div(HW, 256), N), 256, 0, stream>>>(x, label, m, arg, C, HW);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return wsegk_select_kth(m, N, HW, HW / 4, 0, sel, idx_cut, workspace, workspace_bytes, stream);
}

extern "C" int wsegk_minpool_bwd(const float* m, const int32_t* arg, const float* sel, const int32_t* idx_cut,
                                 const float* label, const float* grad_loss, float* grad_x, int N, int C, int H, int W,
                                 wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(m && arg && sel && idx_cut && label && grad_loss && grad_x, WSEGK_ERR_NULL, "minpool_bwd: null pointer");
    WSEGK_REQUIRE(N > 0 && C >= 2 && H > 0 && W > 0, WSEGK_ERR_BAD_SHAPE, "minpool_bwd: bad dims");
    const int HW = H * W;
    minpool_bwd_kernel<<<dim3(ceil_div(HW, 256), N), 256, 0, stream>>>(m, arg, reinterpret_cast<const float4*>(sel), idx_cut,
                                                                        label, grad_loss, grad_x, N, C, HW, HW / 4);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

static int up_dims_ok(const char* who, int N, int C, int h, int w, int H, int W) {
    WSEGK_REQUIRE(N > 0 && C >= 2 && C <= 32 && h > 0 && w > 0 && H >= h && W >= w && static_cast<long long>(H) * W < (1ll << 30),
                  WSEGK_ERR_BAD_SHAPE, "%s: need N > 0, 2 <= C <= 32, 0 < h <= H, 0 < w <= W (got N=%d C=%d %dx%d -> %dx%d)", who,
                  N, C, h, w, H, W);
    return WSEGK_OK;
}

extern "C" int wsegk_gap_mlsm_up_fwd(const float* small, const float* label, float* gap, float* dgap, float* loss, int N,
                                     int C, int h, int w, int H, int W, void* workspace, size_t workspace_bytes,
                                     wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(small && label && gap && dgap && loss && workspace, WSEGK_ERR_NULL, "gap_mlsm_up_fwd: null pointer");
    if (int rc = up_dims_ok("gap_mlsm_up_fwd", N, C, h, w, H, W)) return rc;
    WSEGK_REQUIRE(workspace_bytes >= static_cast<size_t>(N) * C * kGapSplit * sizeof(float), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "gap_mlsm_up_fwd: workspace too small");
    float* partial = static_cast<float*>(workspace);
    gap_up_kernel<<<N * C, 256, static_cast<size_t>(h + w) * sizeof(float), stream>>>(small, partial, up_geom(h, w, H, W));
    gap_mlsm_kernel<<<1, 256, 0, stream>>>(partial, label, gap, dgap, loss, N, C, H * W);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_gap_mlsm_up_bwd(const float* dgap, const float* grad_loss, float* grad_small, int N, int C, int h, int w,
                                     int H, int W, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(dgap && grad_loss && grad_small, WSEGK_ERR_NULL, "gap_mlsm_up_bwd: null pointer");
    if (int rc = up_dims_ok("gap_mlsm_up_bwd", N, C, h, w, H, W)) return rc;
    gap_up_bwd_kernel<<<N * C, 256, static_cast<size_t>(h + w) * sizeof(float), stream>>>(dgap, grad_loss, grad_small,
                                                                                          up_geom(h, w, H, W));
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_minpool_up_fwd(const float* small, const float* label, float* m, int32_t* arg, float* sel,
                                    int32_t* idx_cut, int N, int C, int h, int w, int H, int W, void* workspace,
                                    size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(small && label && m && arg && sel && idx_cut && workspace, WSEGK_ERR_NULL, "minpool_up_fwd: null pointer");
    if (int rc = up_dims_ok("minpool_up_fwd", N, C, h, w, H, W)) return rc;
    WSEGK_REQUIRE(H * W >= 4, WSEGK_ERR_BAD_SHAPE, "minpool_up_fwd: bad dims");
    const int HW = H * W;
    minpool_max_up_kernel<<<dim3(ceil_div(HW, 256), N), 256, 0, stream>>>(small, label, m, arg, C, up_geom(h, w, H, W));
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return wsegk_select_kth(m, N, HW, HW / 4, 0, sel, idx_cut, workspace, workspace_bytes, stream);
}

extern "C" int wsegk_minpool_up_bwd(const float* m, const int32_t* arg, const float* sel, const int32_t* idx_cut,
                                    const float* label, const float* grad_loss, float* grad_small, int N, int C, int h, int w,
                                    int H, int W, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(m && arg && sel && idx_cut && label && grad_loss && grad_small && workspace, WSEGK_ERR_NULL,
                  "minpool_up_bwd: null pointer");
    if (int rc = up_dims_ok("minpool_up_bwd", N, C, h, w, H, W)) return rc;
    WSEGK_REQUIRE(workspace_bytes >= lazy_bwd_workspace_bytes(N, C, H, W), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "minpool_up_bwd: workspace %zu < %zu (WSEGK_OP_LAZY_BWD)", workspace_bytes,
                  lazy_bwd_workspace_bytes(N, C, H, W));
    float* scratch = reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + fullres_workspace_bytes(N, C));
    const int HW = H * W;
    minpool_dense_labelled_kernel<<<dim3(ceil_div(HW, 256), N), 256, 0, stream>>>(
        m, arg, reinterpret_cast<const float4*>(sel), idx_cut, label, grad_loss, scratch, N, C, HW, HW / 4);
    count_launches(1);
    launch_resize_bwd_labelled(scratch, label, grad_small, N * C, h, w, H, W, C, stream);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_er_ecr_fwd(const float* n1, const float* n2, const float* rv1, const float* rv2, float* e1, float* e2,
                                float* sel1, int32_t* cut1, float* sel2, int32_t* cut2, float* losses, int N, int C, int h,
                                int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(n1 && n2 && rv1 && rv2 && e1 && e2 && sel1 && cut1 && sel2 && cut2 && losses && workspace, WSEGK_ERR_NULL,
                  "er_ecr_fwd: null pointer");
    WSEGK_REQUIRE(N > 0 && C >= 2 && C <= kAuxMaxC && h > 0 && w > 0, WSEGK_ERR_BAD_SHAPE, "er_ecr_fwd: need 2 <= C <= %d",
                  kAuxMaxC);
    const int hw = h * w;
    const int k = static_cast<int>(static_cast<double>(C) * hw * 0.2);          // int(21 * hs * ws * 0.2)
    WSEGK_REQUIRE(k >= 1, WSEGK_ERR_BAD_SHAPE, "er_ecr_fwd: map too small");
    const int nblk = ceil_div(hw, 128);
    const size_t part_bytes = align_up(static_cast<size_t>(N) * nblk * sizeof(float), 256);
    WSEGK_REQUIRE(workspace_bytes >= part_bytes + select_ws_bytes(N), WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "er_ecr_fwd: workspace %zu < %zu", workspace_bytes, part_bytes + select_ws_bytes(N));
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, WSEGK_ERR_BAD_ALIGN,
                  "er_ecr_fwd: workspace must be 256-byte aligned");
    float* part = static_cast<float*>(workspace);
    void* sel_ws = static_cast<uint8_t*>(workspace) + part_bytes;
    er_ecr_fwd_kernel<<<dim3(nblk, N), 128, 0, stream>>>(n1, n2, rv1, rv2, e1, e2, part, C, hw);
    count_launches(1);
    // the two selects run one after the other on the stream, so they can share the scratch
    if (int rc = wsegk_select_kth(e1, N, C * hw, k, 1, sel1, cut1, sel_ws, select_ws_bytes(N), stream)) return rc;
    if (int rc = wsegk_select_kth(e2, N, C * hw, k, 1, sel2, cut2, sel_ws, select_ws_bytes(N), stream)) return rc;
    er_ecr_finalize_kernel<<<1, 256, 0, stream>>>(part, N * nblk, reinterpret_cast<const float4*>(sel1),
                                                  reinterpret_cast<const float4*>(sel2), losses, N, C, hw, k);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_er_ecr_bwd(const float* n1, const float* n2, const float* rv1, const float* rv2, const float* sel1,
                                const int32_t* cut1, const float* sel2, const int32_t* cut2, const float* grad_er,
                                const float* grad_ecr, float* gn1, float* gn2, float* grv1, float* grv2, int N, int C, int h,
                                int w, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(n1 && n2 && rv1 && rv2 && sel1 && cut1 && sel2 && cut2 && grad_er && grad_ecr && gn1 && gn2 && grv1 && grv2,
                  WSEGK_ERR_NULL, "er_ecr_bwd: null pointer");
    WSEGK_REQUIRE(N > 0 && C >= 2 && C <= kAuxMaxC && h > 0 && w > 0, WSEGK_ERR_BAD_SHAPE, "er_ecr_bwd: bad dims");
    const int hw = h * w;
    const int k = static_cast<int>(static_cast<double>(C) * hw * 0.2);
    er_ecr_bwd_kernel<<<dim3(ceil_div(hw, 128), N), 128, 0, stream>>>(
        n1, n2, rv1, rv2, reinterpret_cast<const float4*>(sel1), cut1, reinterpret_cast<const float4*>(sel2), cut2, grad_er,
        grad_ecr, gn1, gn2, grv1, grv2, N, C, hw, k);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}
