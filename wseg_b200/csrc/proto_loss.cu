// K3 — per-class top-k of CAM-confident pixels + CAM-weighted prototypes   (contrast_train.py:199-209 / 229-241)
// K4 — pixel-to-prototype NCE losses, forward and backward fused            (contrast_train.py:244-389)
//
// Selection semantics are canonical (value, then lowest flat index) so results are bit-reproducible; every
// floating-point reduction is done in a fixed order (no float atomics).
//
// Two size regimes.  The literal loss grid (16x16, P = N*256 pixels) is latency-bound: one CTA per class for the top-k,
// 16-pixel CTAs and pairwise in-class ranks for the mining.  Large grids (P > 8192 / 16384) switch to the multi-CTA radix
// select of segselect.cuh (top-k and mining thresholds) and to persistent 64-/32-pixel-tile score and gradient kernels.
// Also here: the one-shot NVLink peer-memory exchanges of the data-parallel path (prototype candidates; per-pixel mining
// statistics for the globally exact mining) and the Philox mining keys.
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "segselect.cuh"

namespace wsegk {

// ================================================================================================
// K3
// ================================================================================================
__device__ __forceinline__ uint32_t ordered_bits(float v) {
    if (v == 0.f) v = 0.f;                       // -0 -> +0 (torch.sort treats them as equal)
    uint32_t u = __float_as_uint(v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unordered_bits(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// larger key == earlier in the canonical order (value desc, index asc); keys are unique
__device__ __forceinline__ unsigned long long topk_key(float v, uint32_t idx) {
    return (static_cast<unsigned long long>(ordered_bits(v)) << 32) | (0xFFFFFFFFu - idx);
}

// in-place bitonic sort of n (power of two) 64-bit keys in shared memory, descending, with an optional payload
template <bool kPayload>
__device__ void bitonic_desc(unsigned long long* keys, uint32_t* pay, int n) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (n >> 1); t += blockDim.x) {
                const int lo = 2 * t - (t & (stride - 1));
                const int hi = lo + stride;
                const bool desc = ((lo & size) == 0);
                const unsigned long long a = keys[lo], b = keys[hi];
                if ((a < b) == desc) {
                    keys[lo] = b; keys[hi] = a;
                    if (kPayload) { const uint32_t x = pay[lo]; pay[lo] = pay[hi]; pay[hi] = x; }
                }
            }
        }
    }
    __syncthreads();
}

constexpr int kTopkThreads = 1024;
constexpr int kTopkMaxK = 4096;

// tail shared by both selection paths: s_keys holds the k selected keys (rest zero) -> sort (value desc, index asc),
// write values and global pixel ids, gather the (un-normalised) features top_fea[c][t][d] = f_proj[n][d][q]
__device__ void topk_emit(unsigned long long* s_keys, int c, int k, int kpad, int pixel_offset,
                          const float* __restrict__ f_proj, float* __restrict__ top_val, int32_t* __restrict__ top_idx,
                          float* __restrict__ top_fea, int D, int HW) {
    bitonic_desc<false>(s_keys, nullptr, kpad);
    for (int t = threadIdx.x; t < k; t += blockDim.x) {
        const unsigned long long key = s_keys[t];
        const uint32_t idx = 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
        top_val[c * k + t] = unordered_bits(static_cast<uint32_t>(key >> 32));     // the value the key was built from
        top_idx[c * k + t] = static_cast<int32_t>(idx) + pixel_offset;
    }
    // eight independent (scattered) loads per thread in flight
    const int total = k * D, step = static_cast<int>(blockDim.x);
    for (int e0 = threadIdx.x; e0 < total; e0 += 8 * step) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + j * step;
            v[j] = 0.f;
            if (e < total) {
                const int t = e / D, d = e - t * D;
                const uint32_t idx = 0xFFFFFFFFu - static_cast<uint32_t>(s_keys[t] & 0xFFFFFFFFull);
                const int n = idx / HW, q = idx - n * HW;
                v[j] = __ldg(f_proj + (static_cast<size_t>(n) * D + d) * HW + q);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + j * step;
            if (e < total) top_fea[static_cast<size_t>(c) * total + e] = v[j];
        }
    }
}

// Small planes (P <= kTopkRegs * 1024): one CTA per class radix-selects the k largest keys over the whole batch plane;
// the keys are built once and live in registers, the select passes only touch shared memory.
constexpr int kTopkRegs = 8;
__global__ void __launch_bounds__(kTopkThreads) proto_topk_kernel(const float* __restrict__ norm_cam,
                                                                  const float* __restrict__ f_proj,
                                                                  int pixel_offset, int k, int kpad,
                                                                  float* __restrict__ top_val,
                                                                  int32_t* __restrict__ top_idx,
                                                                  float* __restrict__ top_fea, int N, int C,
                                                                  int D, int HW) {
    extern __shared__ unsigned long long s_keys[];      // kpad
    __shared__ uint32_t s_hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_need, s_done, s_count;
    const int c = blockIdx.x;
    const int P = N * HW;
    unsigned long long kreg[kTopkRegs];
#pragma unroll
    for (int j = 0; j < kTopkRegs; ++j) {
        const int i = threadIdx.x + j * kTopkThreads;
        if (i < P) {
            const int n = i / HW, q = i - n * HW;
            kreg[j] = topk_key(__ldg(norm_cam + (static_cast<size_t>(n) * C + c) * HW + q), static_cast<uint32_t>(i));
        } else {
            kreg[j] = 0ull;
        }
    }
    // visit every key of the plane: f(key)
    auto for_each_key = [&](auto&& f) {
#pragma unroll
        for (int j = 0; j < kTopkRegs; ++j)
            if (threadIdx.x + j * kTopkThreads < P) f(kreg[j]);
    };
    if (threadIdx.x == 0) { s_prefix = 0ull; s_need = k; s_done = 0; s_count = 0; }
    int shift = 56;
    for (int pass = 0; pass < 8; ++pass) {
        shift = 56 - 8 * pass;
        for (int b = threadIdx.x; b < 256; b += blockDim.x) s_hist[b] = 0;
        __syncthreads();
        const unsigned long long prefix = s_prefix;
        // warp-aggregated histogram: CAM values share their sign/exponent byte, so plain shared-memory atomics would
        // serialise thousands of increments on two or three bins in the first passes
        for_each_key([&](unsigned long long key) {
            const bool act = pass == 0 || (key >> (shift + 8)) == prefix;
            const unsigned am = __ballot_sync(__activemask(), act);
            if (act) {
                const uint32_t bin = static_cast<uint32_t>(key >> shift) & 255u;
                const unsigned peers = __match_any_sync(am, bin);
                if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&s_hist[bin], static_cast<uint32_t>(__popc(peers)));
            }
        });
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp 0 finds the bin holding the need-th largest key: lane l owns bins 255-8l .. 248-8l (descending),
            // an inclusive scan over lanes locates the owner, which then walks its 8 bins
            const int lane = threadIdx.x;
            int local[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { local[j] = static_cast<int>(s_hist[255 - (8 * lane + j)]); sum += local[j]; }
            int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int v = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += v;
            }
            const int need = s_need, excl = incl - sum;
            if (excl < need && need <= incl) {
                int cum = excl;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (cum + local[j] >= need) {
                        const int left = need - cum;
                        s_need = left;
                        s_done = (local[j] == left) ? 1 : 0;
                        s_prefix = (prefix << 8) | static_cast<unsigned long long>(255 - (8 * lane + j));
                        break;
                    }
                    cum += local[j];
                }
            }
        }
        __syncthreads();
        if (s_done) break;
    }
    // every key whose top bits are >= the prefix is selected (exactly k of them)
    const unsigned long long prefix = s_prefix;
    for (int i = threadIdx.x; i < kpad; i += blockDim.x) s_keys[i] = 0ull;
    __syncthreads();
    for_each_key([&](unsigned long long key) {
        if ((key >> shift) >= prefix) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < kpad) s_keys[slot] = key;
        }
    });
    topk_emit(s_keys, c, k, kpad, pixel_offset, f_proj, top_val, top_idx, top_fea, D, HW);
}

// Large planes: the thresholds come from the multi-CTA select (segselect.cuh) over all N*C*HW values at once.  Ascending
// composite key = (~ordered(value)) above the flat pixel id: rank k-1 in ascending order is the k-th best pixel of the
// canonical order (value desc, index asc), and the members are exactly the keys <= that threshold.
struct TopkSrc {
    const float* norm_cam;
    int C, HW, idx_bits;
    long long total;                  // N * C * HW
    static constexpr int kSel = 1;
    struct Elem { float v; };
    __device__ long long count() const { return total; }
    __device__ Elem load(long long e) const { return Elem{__ldg(norm_cam + e)}; }
    __device__ void get(const Elem& x, long long e, int, int& slot, unsigned long long& key) const {
        int c;
        long long i;                                             // flat pixel id n*HW + q
        if (total <= 0x7fffffffll) {                             // 32-bit divisions where they suffice
            const uint32_t e32 = static_cast<uint32_t>(e), plane = static_cast<uint32_t>(C) * HW;
            const uint32_t n = e32 / plane, r = e32 - n * plane;
            const uint32_t cc = r / static_cast<uint32_t>(HW);
            c = static_cast<int>(cc);
            i = static_cast<long long>(n * HW + (r - cc * HW));
        } else {
            const long long plane = static_cast<long long>(C) * HW;
            const long long n = e / plane, r = e - n * plane;
            c = static_cast<int>(r / HW);
            i = n * HW + (r - static_cast<long long>(c) * HW);
        }
        slot = c;
        key = (static_cast<unsigned long long>(static_cast<uint32_t>(~ordered_bits(x.v))) << idx_bits) |
              static_cast<unsigned long long>(i);
    }
};
__global__ void topk_sel_init_kernel(SegSel* __restrict__ state, int C, int k) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) state[c] = SegSel{0ull, k - 1, 1};
}
__global__ void __launch_bounds__(256) topk_compact_kernel(TopkSrc src, const SegSel* __restrict__ state,
                                                           int* __restrict__ count, unsigned long long* __restrict__ cand,
                                                           int k) {
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long e = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; e < src.total; e += stride) {
        int c;
        unsigned long long key;
        src.get(src.load(e), e, 0, c, key);
        if (key <= state[c].prefix) {
            const int slot = atomicAdd(count + c, 1);
            if (slot < k) cand[static_cast<size_t>(c) * k + slot] = key;
        }
    }
}
__global__ void __launch_bounds__(kTopkThreads) topk_tail_kernel(const unsigned long long* __restrict__ cand,
                                                                 const float* __restrict__ f_proj, int pixel_offset,
                                                                 int k, int kpad, int idx_bits,
                                                                 float* __restrict__ top_val,
                                                                 int32_t* __restrict__ top_idx,
                                                                 float* __restrict__ top_fea, int D, int HW) {
    extern __shared__ unsigned long long s_keys[];      // kpad
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < kpad; i += blockDim.x) {
        unsigned long long key = 0ull;
        if (i < k) {
            const unsigned long long a = cand[static_cast<size_t>(c) * k + i];
            const uint32_t ordv = ~static_cast<uint32_t>(a >> idx_bits);
            const uint32_t idx = static_cast<uint32_t>(a & ((1ull << idx_bits) - 1ull));
            key = (static_cast<unsigned long long>(ordv) << 32) | (0xFFFFFFFFu - idx);
        }
        s_keys[i] = key;
    }
    topk_emit(s_keys, c, k, kpad, pixel_offset, f_proj, top_val, top_idx, top_fea, D, HW);
}

struct TopkLayout { size_t off_state, off_count, off_hist, off_cand, clear_bytes, total; };
static TopkLayout topk_layout(int C) {
    TopkLayout L{};
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.off_state = take(static_cast<size_t>(C) * sizeof(SegSel));
    L.off_count = take(static_cast<size_t>(C) * sizeof(int));
    L.off_hist = take(static_cast<size_t>(C) * kSegBins * sizeof(uint32_t));
    L.clear_bytes = off;
    L.off_cand = take(static_cast<size_t>(C) * kTopkMaxK * sizeof(unsigned long long));
    L.total = align_up(off, 1024);
    return L;
}

// Where the candidates of all ranks live: either three plain tensors ([C][n_cand], [C][n_cand], [C][n_cand][D], what an
// all-gather of the packed messages is unpacked into) or the rank-major packed exchange buffer that the peers' push
// kernels wrote straight into this GPU's memory over NVLink ([R][C][k][D+2]: value, index bits, D features).
struct CandSrc {
    const float* val;
    const int32_t* idx;
    const float* fea;
    const float* packed;
    int n_cand, C, k, D;
    __device__ __forceinline__ size_t row(int c, int j) const {
        const int r = j / k, t = j - r * k;
        return (static_cast<size_t>(r * C + c) * k + t) * (D + 2);
    }
    __device__ __forceinline__ float value(int c, int j) const { return packed ? packed[row(c, j)] : val[c * n_cand + j]; }
    __device__ __forceinline__ uint32_t index(int c, int j) const {
        return packed ? __float_as_uint(packed[row(c, j) + 1]) : static_cast<uint32_t>(idx[c * n_cand + j]);
    }
    __device__ __forceinline__ float feature(int c, int j, int d) const {
        return packed ? packed[row(c, j) + 2 + d] : fea[(static_cast<size_t>(c) * n_cand + j) * D + d];
    }
};

__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// one CTA per class: global top-k among n_cand candidates, prototype = sum(val*fea)/sum(val), L2-normalised.
// wait_flags != nullptr: first wait until every rank's push of step `*wait_step` has landed (flags written by the peers).
__global__ void __launch_bounds__(1024) proto_finalize_kernel(CandSrc src, int npad, float* __restrict__ proto,
                                                              int32_t* __restrict__ sel_idx, const uint32_t* wait_flags,
                                                              const uint32_t* wait_step, int n_ranks) {
    extern __shared__ unsigned long long s_buf[];
    unsigned long long* s_keys = s_buf;                                   // npad
    uint32_t* s_pos = reinterpret_cast<uint32_t*>(s_buf + npad);          // npad
    __shared__ float s_red[32];
    __shared__ float s_scale;
    const int c = blockIdx.x;
    const int n_cand = src.n_cand, k = src.k, D = src.D;
    if (wait_flags != nullptr) {
        if (threadIdx.x < n_ranks) {
            const uint32_t want = *wait_step;
            unsigned long long spins = 0;
            while (static_cast<int32_t>(ld_acquire_sys(wait_flags + threadIdx.x) - want) < 0) {
                if (++spins > (1ull << 33)) __trap();          // a peer died: fail the launch instead of hanging
                __nanosleep(64);
            }
        }
        __syncthreads();
        // the parity half of the exchange buffer that belongs to this step
        src.packed += static_cast<size_t>(*wait_step & 1u) * 2 * n_ranks * src.C * k * (D + 2);
    }
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        if (i < n_cand) {
            s_keys[i] = topk_key(src.value(c, i), src.index(c, i));
            s_pos[i] = i;
        } else {
            s_keys[i] = 0ull;
            s_pos[i] = 0;
        }
    }
    bitonic_desc<true>(s_keys, s_pos, npad);
    for (int t = threadIdx.x; t < k; t += blockDim.x)
        sel_idx[c * k + t] = static_cast<int32_t>(0xFFFFFFFFu - static_cast<uint32_t>(s_keys[t] & 0xFFFFFFFFull));
    // selected weights into shared memory (one load per thread), then a fixed-order sum by thread 0
    float* s_w = reinterpret_cast<float*>(s_keys);             // the keys are no longer needed: reuse as k floats
    __syncthreads();
    for (int t = threadIdx.x; t < k; t += blockDim.x) s_w[t] = src.value(c, s_pos[t]);
    __syncthreads();
    if (threadIdx.x == 0) {
        float wsum = 0.f;
        for (int t = 0; t < k; ++t) wsum += s_w[t];
        s_scale = wsum;
    }
    __syncthreads();
    const float wsum = s_scale;
    float sq = 0.f;
    float mine[4] = {0.f, 0.f, 0.f, 0.f};          // D <= 4 * blockDim
    const int Dp = (D + 31) & ~31;
    if (Dp <= 512 && blockDim.x == 1024) {
        // thread (part, d): the k candidates are split into 1024/Dp contiguous parts so all threads gather; the parts
        // are combined in a fixed order
        __shared__ float s_acc[1024];
        const int nparts = 1024 / Dp;
        const int part = threadIdx.x / Dp, d = threadIdx.x - part * Dp;
        const int per = (k + nparts - 1) / nparts;
        if (part < nparts && d < D) {
            const int t_end = min(k, (part + 1) * per);
            float acc = 0.f;
            for (int t0 = part * per; t0 < t_end; t0 += 8) {  // eight independent loads in flight, fixed order of adds
                float fv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) fv[j] = (t0 + j < t_end) ? src.feature(c, s_pos[t0 + j], d) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (t0 + j < t_end) acc = fmaf(s_w[t0 + j], fv[j], acc);
            }
            s_acc[part * Dp + d] = acc;
        }
        __syncthreads();
        if (threadIdx.x < D) {
            float acc = 0.f;
            for (int q = 0; q < nparts; ++q) acc += s_acc[q * Dp + threadIdx.x];
            acc /= wsum;
            mine[0] = acc;
            sq = acc * acc;
        }
    } else {
        int slot = 0;
        for (int d = threadIdx.x; d < D; d += blockDim.x, ++slot) {
            float acc = 0.f;
            for (int t0 = 0; t0 < k; t0 += 8) {               // eight independent loads in flight, fixed order of adds
                float fv[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) fv[j] = (t0 + j < k) ? src.feature(c, s_pos[t0 + j], d) : 0.f;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (t0 + j < k) acc = fmaf(s_w[t0 + j], fv[j], acc);
            }
            acc /= wsum;
            mine[slot] = acc;
            sq = fmaf(acc, acc, sq);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int wv = 0; wv < (blockDim.x >> 5); ++wv) tot += s_red[wv];
        s_scale = fmaxf(sqrtf(tot), 1e-12f);          // F.normalize eps
    }
    __syncthreads();
    const float nrm = s_scale;
    int slot = 0;
    for (int d = threadIdx.x; d < D; d += blockDim.x, ++slot) proto[c * D + d] = mine[slot] / nrm;
}

// ---- one-shot candidate exchange over NVLink peer memory ------------------------------------------------
// Every rank owns the same symmetric buffer: candidates [2 parities][2 views][R ranks][C][k][D+2] floats, then flags
// [2 views][R] (the step number of the last complete push of that rank).  The push kernel packs this rank's local top-k
// and stores it into slot (parity, view, my rank) of EVERY peer's buffer (plain stores to mapped peer addresses), fences,
// and the last CTA to finish releases flag[view][my rank] = step on every peer.  proto_finalize then only waits for R flags
// in its own memory.  No NCCL kernel, no host involvement, graph-capturable; the step number lives on the device.
// The wait for the peers' pushes, in a kernel of ONE warp that fits beside anything (32 threads, < 32 registers): while
// the 21 CTAs x 1024 threads of proto_finalize_kernel did the spinning themselves they held 21 SMs for as long as the slowest
// peer was behind, and the persistent PCM kernels (one whole-SM CTA on 144 of the 148 SMs) could not be placed until the
// spin ended — the step lost the ranks' skew (~40 us) from 2 GPUs on (SCALE_r01).
__global__ void __launch_bounds__(32) proto_wait_flags_kernel(const uint32_t* wait_flags, const uint32_t* wait_step,
                                                              int n_ranks, unsigned long long* stamp) {
    if (static_cast<int>(threadIdx.x) < n_ranks) {
        const uint32_t want = *wait_step;
        unsigned long long t0 = 0;
        if (stamp != nullptr && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t0));
        unsigned long long spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys(wait_flags + threadIdx.x) - want) < 0) {
            if (++spins > (1ull << 33)) __trap();          // a peer died: fail the launch instead of hanging
            __nanosleep(64);
        }
        if (stamp != nullptr) {
            // diagnostics (wsegk_debug_exchange_stamps): [0] when this rank started waiting, [1 + r] when rank r's flag was seen
            unsigned long long t1;
            asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t1));
            if (threadIdx.x == 0) stamp[0] = t0;
            stamp[1 + threadIdx.x] = t1;
        }
    }
}

constexpr int kMaxRanks = 16;
struct PushPeers {
    float* cand[kMaxRanks];
    uint32_t* flags[kMaxRanks];
};
constexpr int kPushChunks = 16;

__global__ void __launch_bounds__(256) proto_push_kernel(const float* __restrict__ top_val, const int32_t* __restrict__ top_idx,
                                                         const float* __restrict__ top_fea, PushPeers peers, int R, int rank,
                                                         int view, int C, int k, int D, uint32_t* __restrict__ state) {
    // state[view*2 + 0] = last completed step, state[view*2 + 1] = CTAs done in this launch
    const uint32_t step = state[view * 2] + 1u;
    const int peer = blockIdx.y;
    const size_t slot_floats = static_cast<size_t>(C) * k * (D + 2);
    float* dst = peers.cand[peer] + ((static_cast<size_t>(step & 1u) * 2 + view) * R + rank) * slot_floats;
    const int rows = C * k, pairs = (D + 2) >> 1;                  // D even: a row is (D+2)/2 float2
    const int per = ceil_div_dev(rows, static_cast<int>(gridDim.x));
    const int r0 = blockIdx.x * per, r1 = min(rows, r0 + per);
    // eight elements per thread and pass, all loads before the first (remote) store: rolled, every iteration paid one L2
    // round trip for its load before its store could leave (r02 timeline at 8 GPUs: ~40 us between the top-k and the
    // first flag)
    const int e_end = r1 * pairs;
    for (int e0 = r0 * pairs + threadIdx.x; e0 < e_end; e0 += 8 * blockDim.x) {
        float2 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + j * static_cast<int>(blockDim.x);
            v[j] = make_float2(0.f, 0.f);
            if (e < e_end) {
                const int row = e / pairs, u = e - row * pairs;
                if (u == 0) {
                    v[j].x = __ldg(top_val + row);
                    v[j].y = __int_as_float(__ldg(top_idx + row));
                } else {
                    const float2 f2 = __ldg(reinterpret_cast<const float2*>(top_fea + static_cast<size_t>(row) * D) + (u - 1));
                    v[j] = f2;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int e = e0 + j * static_cast<int>(blockDim.x);
            if (e < e_end) {
                const int row = e / pairs, u = e - row * pairs;
                reinterpret_cast<float2*>(dst + static_cast<size_t>(row) * (D + 2))[u] = v[j];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned done = atomicAdd(&state[view * 2 + 1], 1u);
        if (done == total - 1) {                                   // last CTA: every store of this launch is fenced
            state[view * 2 + 1] = 0u;
            state[view * 2] = step;
            __threadfence_system();
            for (int p = 0; p < R; ++p) st_release_sys(peers.flags[p] + view * R + rank, step);
        }
    }
}

size_t proto_workspace_bytes(int, int C, int, int) { return topk_layout(C).total; }

// ================================================================================================
// K4
// ================================================================================================
constexpr int kNceMaxC = 24;
constexpr int kNceThreads = 128;
constexpr float kInvT = 10.f;          // 1 / temperature 0.1

struct NceView {
    const float* x;            // [N,D,HW]
    const float* proto_own;    // [C,D]
    const float* proto_oth;
    const int32_t* pl_own;     // [P]
    const int32_t* pl_oth;
    const uint32_t* rkey;      // [P]
    float* grad_x;             // [N,D,HW]
    int32_t* mined;            // [P] or null
    // workspace (per view)
    float* inv_norm;           // [P]
    float* sim;                // [P]
    float* ell;                // [P]  unweighted intra loss of the pixel
    uint32_t* setmask;         // [P]  classes at ranks 3..12 of q.proto_own
    float* s_own;              // [C][P]
    float* s_oth;              // [C][P]
    int* rank_sim;             // [P]
    int* rank_rnd;             // [P]
    int* cls_count;            // [32]
    SegSel* sel_state;         // [3][kNceMaxC] thresholds of the large-P mining path (keys at ranks half, kk-half, kk)
    uint32_t* sel_hist;        // [3][kNceMaxC][kSegBins]
    float* partial;            // [nblk][4]: cross, cross2, intra(weighted), unused
};
struct NceParams {
    NceView v[2];
    int N, C, D, HW, P;
    int nblk, nblk_grad;       // tiles of the pixel kernel (partials 0,1) and of the gradient kernel (partial 2)
    int idx_bits;              // > 0: mining by thresholds (sel_state) instead of pairwise ranks
    int idx_off;               // index of this rank's first pixel in the keys of the thresholds (global mining: rank * P)
    float intra_scale;         // 1, or the number of ranks when the mining is global (see nce_mine_global_kernel)
    const float* grad_scale;
    float* losses;
};

// Work split of kernels A and C: a CTA owns kNcePix pixels; each pixel is served by kNceGroups threads.  The
// (prototype set, class) pairs form 48 "slots" (0..23 own prototypes, 24..47 the other view's); thread group g owns
// slots 6g..6g+5 in A (dot products over all D) and D/8-strided channels in C (gradient rows).  Every dot product is
// still accumulated sequentially over d by ONE thread, so values do not depend on the launch geometry.
constexpr int kNcePix = 16;
constexpr int kNceGroups = 8;
constexpr int kNceSlots = 48;
constexpr int kSlotHalf = 24;
static_assert(kNcePix * kNceGroups == kNceThreads, "thread mapping");

// 48 x D prototype values into shared memory, eight independent loads per thread in flight (a rolled loop of
// dependent load->store pairs cost 48 L2 latencies, ~20 us, in both kernels)
__device__ __forceinline__ void nce_load_protos(const NceView& V, float* s_proto, int C, int D, int pstride) {
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        for (int s0 = 0; s0 < kNceSlots; s0 += 8) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int slot = s0 + j, c = slot % kSlotHalf;
                const float* src = slot < kSlotHalf ? V.proto_own : V.proto_oth;
                v[j] = c < C ? __ldg(src + c * D + d) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s_proto[(s0 + j) * pstride + d] = v[j];
        }
    }
}

// x tile of kNcePix consecutive loss pixels: every thread issues its D/8 loads at once (one memory latency instead of
// D dependent ones); for a fixed channel the pixels are contiguous in memory (they may straddle two images)
__device__ __forceinline__ void nce_load_x(const NceView& V, float* s_x, int p0, int D, int HW, int P) {
    const int pix = threadIdx.x & (kNcePix - 1), d0 = threadIdx.x / kNcePix;      // thread (pix, d0) takes d0, d0+8, ...
    const int p = p0 + pix;
    const bool live = p < P;
    const int n = live ? p / HW : 0, q = live ? p - n * HW : 0;
    const float* src = V.x + static_cast<size_t>(n) * D * HW + q;
    for (int d = d0; d < D; d += 8 * kNceGroups) {                                 // D % 32 == 0
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int dd = d + j * kNceGroups;
            v[j] = (live && dd < D) ? __ldg(src + static_cast<size_t>(dd) * HW) : 0.f;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int dd = d + j * kNceGroups;
            if (dd < D) s_x[dd * kNcePix + pix] = v[j];
        }
    }
}

// x tile of NPIX (= 64) consecutive loss pixels as [D][NPIX]; float4 loads when the layout allows (a quad of pixels
// then never straddles two images), eight independent loads per thread in flight either way
template <int NPIX>
__device__ __forceinline__ void nce_tile_load_x(const NceView& V, float* s_x, int p0, int D, int HW, int P) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(V.x) & 15) == 0) {
        constexpr int Q = NPIX / 4;
        const int quad = tid % Q, d0 = tid / Q, dstep = nthr / Q;
        const int p = p0 + 4 * quad;
        const bool live = p < P;                               // P = N*HW is a multiple of 4 too
        const int n = live ? p / HW : 0, q = live ? p - n * HW : 0;
        const float4* src = reinterpret_cast<const float4*>(V.x + static_cast<size_t>(n) * D * HW + q);
        const size_t row = static_cast<size_t>(HW >> 2);
        for (int d = d0; d < D; d += 8 * dstep) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int dd = d + j * dstep;
                v[j] = (live && dd < D) ? __ldg(src + dd * row) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int dd = d + j * dstep;
                if (dd < D) *reinterpret_cast<float4*>(s_x + dd * NPIX + 4 * quad) = v[j];
            }
        }
    } else {
        const int pix = tid % NPIX, d0 = tid / NPIX, dstep = nthr / NPIX;
        const int p = p0 + pix;
        const bool live = p < P;
        const int n = live ? p / HW : 0, q = live ? p - n * HW : 0;
        const float* src = V.x + static_cast<size_t>(n) * D * HW + q;
        for (int d = d0; d < D; d += 8 * dstep) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int dd = d + j * dstep;
                v[j] = (live && dd < D) ? __ldg(src + static_cast<size_t>(dd) * HW) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int dd = d + j * dstep;
                if (dd < D) s_x[dd * NPIX + pix] = v[j];
            }
        }
    }
}

// ---- A: per-pixel similarities and unweighted loss terms ----------------------------------------
__global__ void __launch_bounds__(kNceThreads) nce_pixel_kernel(NceParams prm) {
    extern __shared__ float s_dyn[];
    const NceView& V = prm.v[blockIdx.y];
    const int C = prm.C, D = prm.D, HW = prm.HW, P = prm.P;
    float* s_proto = s_dyn;                                   // [48][D+1]
    float* s_sc = s_dyn + kNceSlots * (D + 1);                // [kNcePix][49]
    float* s_x = s_sc + kNcePix * 49 + kNcePix * kNceGroups * 4;   // [D][kNcePix] features of the CTA's pixels
    nce_load_protos(V, s_proto, C, D, D + 1);
    nce_load_x(V, s_x, blockIdx.x * kNcePix, D, HW, P);
    __syncthreads();
    const int pix = threadIdx.x & (kNcePix - 1), grp = threadIdx.x / kNcePix;
    const int p = blockIdx.x * kNcePix + pix;
    float inv = 0.f;
    if (p < P) {
        const float* pr = s_proto + (grp * 6) * (D + 1);
        float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float ss = 0.f;
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            const float xv = s_x[d * kNcePix + pix];
            ss = fmaf(xv, xv, ss);
#pragma unroll
            for (int j = 0; j < 6; ++j) acc[j] = fmaf(xv, pr[j * (D + 1) + d], acc[j]);
        }
        inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);          // F.normalize
#pragma unroll
        for (int j = 0; j < 6; ++j) s_sc[pix * 49 + grp * 6 + j] = acc[j] * inv;
    }
    __syncthreads();
    // per-pixel scalars, spread over the pixel's 8 threads: thread (pix, grp) owns classes grp, grp+8, grp+16
    float* s_part = s_sc + kNcePix * 49;                      // [kNcePix][kNceGroups][4]
    if (p < P) {
        float z_own = 0.f, z_oth = 0.f, a2 = 0.f;
        uint32_t mask = 0u;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = grp + kNceGroups * j;
            if (c < C) {
                const float own_c = s_sc[pix * 49 + c], oth_c = s_sc[pix * 49 + kSlotHalf + c];
                V.s_own[static_cast<size_t>(c) * P + p] = own_c;
                V.s_oth[static_cast<size_t>(c) * P + p] = oth_c;
                const float e_own = expf(own_c * kInvT);
                z_own += e_own;
                z_oth += expf(oth_c * kInvT);
                // rank of c among the pixel's own-prototype scores (value desc, class asc); ranks 3..12 are semi-hard
                int rank = 0;
                for (int e = 0; e < C; ++e) {
                    const float own_e = s_sc[pix * 49 + e];
                    rank += (own_e > own_c || (own_e == own_c && e < c)) ? 1 : 0;
                }
                if (rank >= 3 && rank < 13) { mask |= (1u << c); a2 += e_own; }
            }
        }
        float* dstp = s_part + (pix * kNceGroups + grp) * 4;
        dstp[0] = z_own; dstp[1] = z_oth; dstp[2] = a2; dstp[3] = __uint_as_float(mask);
    }
    __syncthreads();
    float cross = 0.f, cross2 = 0.f;
    if (grp == 0 && p < P) {
        float z_own = 0.f, z_oth = 0.f, a2 = 0.f;
        uint32_t mask = 0u;
#pragma unroll
        for (int g = 0; g < kNceGroups; ++g) {                 // fixed order
            const float* sp = s_part + (pix * kNceGroups + g) * 4;
            z_own += sp[0]; z_oth += sp[1]; a2 += sp[2]; mask |= __float_as_uint(sp[3]);
        }
        const int lo = V.pl_own[p], lt = V.pl_oth[p];
        const float pos_own = s_sc[pix * 49 + lo], pos_oth_lo = s_sc[pix * 49 + kSlotHalf + lo];
        const float pos_own_lt = s_sc[pix * 49 + lt];
        // cross: other view's prototypes, own label (:261-267); cross2: own prototypes, other view's label (:272-279)
        cross = logf(z_oth) - pos_oth_lo * kInvT;
        cross2 = logf(z_own) - pos_own_lt * kInvT;
        // intra (:285-299): positive + the prototypes at ranks 3..12 of q.proto_own
        a2 += expf(pos_own * kInvT);
        V.setmask[p] = mask;
        V.ell[p] = logf(a2) - pos_own * kInvT;
        V.sim[p] = (pos_own + 1.f) * 0.5f;
        V.inv_norm[p] = inv;
        atomicAdd(&V.cls_count[lo], 1);
    }
    // deterministic block partials: the kNcePix scalar threads are the first lanes of warp 0
    if (threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cross += __shfl_xor_sync(0xffffffffu, cross, o);
            cross2 += __shfl_xor_sync(0xffffffffu, cross2, o);
        }
        if (threadIdx.x == 0) {
            V.partial[blockIdx.x * 4 + 0] = cross;
            V.partial[blockIdx.x * 4 + 1] = cross2;
        }
    }
}

// ---- A (large P): 64 pixels per CTA, 256 threads, 4 pixels x 3 slots per thread ------------------------------------
// The prototypes are kept transposed ([d][68]: slot 3g+j at column 4g+j) so a thread reads its three as one float4, and its
// four pixels of s_x[d][64] as another: 2 shared-memory instructions per 12 FMAs instead of 7 per 6 in the 16-pixel kernel,
// with 16 warps per SM (two 89 KB CTAs) to hide their latency.  Each dot product is still accumulated over d in order by one
// thread: same values.
constexpr int kTilePix = 64;
constexpr int kGradTilePix = 32;       // 256 threads, two CTAs per SM
constexpr int kTileProtoStride = 68;
constexpr int kPixTileThreads = 256;
__global__ void __launch_bounds__(kPixTileThreads) nce_pixel_tile_kernel(NceParams prm) {
    extern __shared__ __align__(16) float s_dyn[];
    const NceView& V = prm.v[blockIdx.y];
    const int C = prm.C, D = prm.D, HW = prm.HW, P = prm.P;
    float* s_proto = s_dyn;                                    // [D][68]
    float* s_x = s_proto + D * kTileProtoStride;               // [D][64]
    float* s_sc = s_x + D * kTilePix;                          // [64][49]
    float* s_part = s_sc + kTilePix * 49;                      // [64][8][4]
    float* s_inv = s_part + kTilePix * kNceGroups * 4;         // [64]
    __shared__ int s_cls[kNceMaxC];                            // pixels per pseudo-class seen by this CTA
    __shared__ float s_cross[2][2];
    if (threadIdx.x < kNceMaxC) s_cls[threadIdx.x] = 0;
    // the CTA is persistent: the prototypes are staged once (all 48 loads of a thread in flight at once)
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float v[kNceSlots];
#pragma unroll
        for (int slot = 0; slot < kNceSlots; ++slot) {
            const int c = slot % kSlotHalf;
            const float* src = slot < kSlotHalf ? V.proto_own : V.proto_oth;
            v[slot] = c < C ? __ldg(src + c * D + d) : 0.f;
        }
#pragma unroll
        for (int slot = 0; slot < kNceSlots; ++slot)
            s_proto[d * kTileProtoStride + 4 * (slot / 3) + slot % 3] = v[slot];
    }
    // scalar-phase mapping: the 256 threads serve two 16-pixel groups at a time, 8 threads (class groups) per pixel
    const int half = threadIdx.x >> 7, t7 = threadIdx.x & 127;
    const int grp = t7 >> 4;
    for (int tile = blockIdx.x; tile < prm.nblk; tile += gridDim.x) {
    const int p0 = tile * kTilePix;
    nce_tile_load_x<kTilePix>(V, s_x, p0, D, HW, P);
    __syncthreads();
    const int pg = threadIdx.x & 15, sg = threadIdx.x >> 4;           // pixels 4pg..4pg+3, slots 3sg..3sg+2
    float acc[4][3];
    {
        float ss[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[i][j] = 0.f;
        const float* xr = s_x + 4 * pg;
        const float* pr = s_proto + 4 * sg;
        const bool norm_warp = threadIdx.x < 32;                      // warp 0 also accumulates |x|^2 (lanes 0-15 publish it)
#pragma unroll 4
        for (int d = 0; d < D; ++d) {
            const float4 xv = *reinterpret_cast<const float4*>(xr + d * kTilePix);
            const float4 pa = *reinterpret_cast<const float4*>(pr + d * kTileProtoStride);
            const float x[4] = {xv.x, xv.y, xv.z, xv.w};
            const float pv[3] = {pa.x, pa.y, pa.z};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (norm_warp) ss[i] = fmaf(x[i], x[i], ss[i]);
#pragma unroll
                for (int j = 0; j < 3; ++j) acc[i][j] = fmaf(x[i], pv[j], acc[i][j]);
            }
        }
        if (sg == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) s_inv[4 * pg + i] = 1.f / fmaxf(sqrtf(ss[i]), 1e-12f);      // F.normalize
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float inv = s_inv[4 * pg + i];
#pragma unroll
        for (int j = 0; j < 3; ++j) s_sc[(4 * pg + i) * 49 + 3 * sg + j] = acc[i][j] * inv;
    }
    __syncthreads();
    // per-pixel scalars: thread (pix, grp) owns classes grp, grp+8, grp+16 of its pixel
    for (int it = 0; it < kTilePix / 32; ++it) {
        const int pix = (t7 & 15) + 16 * (2 * it + half), p = p0 + pix;
        if (p < P) {
            float z_own = 0.f, z_oth = 0.f, a2 = 0.f;
            uint32_t mask = 0u;
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int c = grp + kNceGroups * j;
                if (c < C) {
                    const float own_c = s_sc[pix * 49 + c], oth_c = s_sc[pix * 49 + kSlotHalf + c];
                    V.s_own[static_cast<size_t>(c) * P + p] = own_c;
                    V.s_oth[static_cast<size_t>(c) * P + p] = oth_c;
                    const float e_own = expf(own_c * kInvT);
                    z_own += e_own;
                    z_oth += expf(oth_c * kInvT);
                    int rank = 0;
                    for (int e = 0; e < C; ++e) {
                        const float own_e = s_sc[pix * 49 + e];
                        rank += (own_e > own_c || (own_e == own_c && e < c)) ? 1 : 0;
                    }
                    if (rank >= 3 && rank < 13) { mask |= (1u << c); a2 += e_own; }
                }
            }
            float* dstp = s_part + grp * 4 * kTilePix + pix;       // [grp][4][64]: conflict-free
            dstp[0] = z_own; dstp[kTilePix] = z_oth; dstp[2 * kTilePix] = a2; dstp[3 * kTilePix] = __uint_as_float(mask);
        }
    }
    __syncthreads();
    float cross = 0.f, cross2 = 0.f;
    if (grp == 0) {
        for (int it = 0; it < kTilePix / 32; ++it) {
            const int pix = t7 + 16 * (2 * it + half), p = p0 + pix;
            if (p >= P) continue;
            float z_own = 0.f, z_oth = 0.f, a2 = 0.f;
            uint32_t mask = 0u;
#pragma unroll
            for (int g = 0; g < kNceGroups; ++g) {                 // fixed order
                const float* sp = s_part + g * 4 * kTilePix + pix;
                z_own += sp[0]; z_oth += sp[kTilePix]; a2 += sp[2 * kTilePix]; mask |= __float_as_uint(sp[3 * kTilePix]);
            }
            const int lo = V.pl_own[p], lt = V.pl_oth[p];
            const float pos_own = s_sc[pix * 49 + lo], pos_oth_lo = s_sc[pix * 49 + kSlotHalf + lo];
            const float pos_own_lt = s_sc[pix * 49 + lt];
            cross += logf(z_oth) - pos_oth_lo * kInvT;
            cross2 += logf(z_own) - pos_own_lt * kInvT;
            a2 += expf(pos_own * kInvT);
            V.setmask[p] = mask;
            V.ell[p] = logf(a2) - pos_own * kInvT;
            V.sim[p] = (pos_own + 1.f) * 0.5f;
            V.inv_norm[p] = s_inv[pix];
            atomicAdd(&s_cls[lo], 1);
        }
    }
    // the pixel threads are lanes 0-15 of warps 0 and 4: a shuffle tree in each, then a fixed-order add
    if (t7 < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            cross += __shfl_xor_sync(0xffffffffu, cross, o);
            cross2 += __shfl_xor_sync(0xffffffffu, cross2, o);
        }
        if (t7 == 0) { s_cross[half][0] = cross; s_cross[half][1] = cross2; }
    }
    __syncthreads();                                           // also: s_x / s_sc / s_part are rewritten by the next tile
    if (threadIdx.x == 0) {
        V.partial[tile * 4 + 0] = s_cross[0][0] + s_cross[1][0];
        V.partial[tile * 4 + 1] = s_cross[0][1] + s_cross[1][1];
    }
    }
    if (threadIdx.x < C && s_cls[threadIdx.x] > 0) atomicAdd(&V.cls_count[threadIdx.x], s_cls[threadIdx.x]);
}

// ---- B: ranks inside each pseudo-class (hard-pixel mining, :302-322) -----------------------------
// rank_sim[p] = #{p' same class : (sim', p') < (sim, p)},  rank_rnd likewise on (key, p).  Integer atomics only.
constexpr int kMineThreads = 256;
constexpr int kMineChunk = 256;
__global__ void __launch_bounds__(kMineThreads) nce_mine_kernel(NceParams prm) {
    const NceView& V = prm.v[blockIdx.z];
    const int P = prm.P;
    __shared__ int s_cls[kMineChunk];
    __shared__ float s_sim[kMineChunk];
    __shared__ uint32_t s_key[kMineChunk];
    const int base = blockIdx.y * kMineChunk;
    for (int i = threadIdx.x; i < kMineChunk; i += blockDim.x) {
        const int pp = base + i;
        s_cls[i] = pp < P ? V.pl_own[pp] : -1;
        s_sim[i] = pp < P ? V.sim[pp] : 0.f;
        s_key[i] = pp < P ? V.rkey[pp] : 0u;
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int cls = V.pl_own[p];
    const float sim = V.sim[p];
    const uint32_t key = V.rkey[p];
    int rs = 0, rr = 0;
    const int lim = min(kMineChunk, P - base);
    for (int i = 0; i < lim; ++i) {
        if (s_cls[i] == cls) {
            const int pp = base + i;
            rs += (s_sim[i] < sim || (s_sim[i] == sim && pp < p)) ? 1 : 0;
            rr += (s_key[i] < key || (s_key[i] == key && pp < p)) ? 1 : 0;
        }
    }
    if (rs) atomicAdd(&V.rank_sim[p], rs);
    if (rr) atomicAdd(&V.rank_rnd[p], rr);
}

// ---- B (data parallel, GLOBAL mining) --------------------------------------------------------
// The reference mines over the whole batch (its DataParallel gathers everything on one GPU, contrast_train.py:64,108).
// With the batch sharded over R ranks, every rank pushes (sim, key, pseudo label) of its P pixels and both views into every
// peer's symmetric buffer over NVLink ([2 parities][R][view][sim|key|label][P] words + R flags; the same one-shot protocol as
// the prototype candidates), then ranks its own pixels against ALL R*P pixels of their class in the canonical order
// (value, global index = rank*P + p): the mined sets, class sizes and class count are exactly the single-process ones.
// Each rank's intra loss is R x its share of the global one, so that the mean over ranks (what DDP does to the
// gradients) is the single-process loss.
struct MinePeers {
    uint32_t* base[kMaxRanks];
    uint32_t* flags[kMaxRanks];
};
__global__ void __launch_bounds__(256) mine_push_kernel(NceParams prm, MinePeers peers, int R, int rank,
                                                        uint32_t* __restrict__ state) {
    const uint32_t step = state[0] + 1u;
    const int peer = blockIdx.y, P = prm.P;
    const size_t words = static_cast<size_t>(6) * P;
    uint32_t* dst = peers.base[peer] + (static_cast<size_t>(step & 1u) * R + rank) * words;
    for (size_t e = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; e < words;
         e += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int v = static_cast<int>(e / (3 * static_cast<size_t>(P)));
        const int r = static_cast<int>(e - static_cast<size_t>(v) * 3 * P);
        const int a = r / P, i = r - a * P;
        const NceView& V = prm.v[v];
        dst[e] = a == 0 ? __float_as_uint(V.sim[i]) : (a == 1 ? V.rkey[i] : static_cast<uint32_t>(V.pl_own[i]));
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned total = gridDim.x * gridDim.y;
        const unsigned done = atomicAdd(&state[1], 1u);
        if (done == total - 1) {                                   // last CTA: every store of this launch is fenced
            state[1] = 0u;
            state[0] = step;
            __threadfence_system();
            for (int p = 0; p < R; ++p) st_release_sys(peers.flags[p] + rank, step);
        }
    }
}
__device__ __forceinline__ void mine_wait(const uint32_t* flags, const uint32_t* state, int R) {
    if (threadIdx.x < R) {
        const uint32_t want = state[0];
        unsigned long long spins = 0;
        while (static_cast<int32_t>(ld_acquire_sys(flags + threadIdx.x) - want) < 0) {
            if (++spins > (1ull << 33)) __trap();              // a peer died: fail the launch instead of hanging
            __nanosleep(64);
        }
    }
    __syncthreads();
}
// global class sizes (one CTA per view): replaces the local counts nce_pixel_kernel accumulated
__global__ void __launch_bounds__(1024) mine_global_count_kernel(NceParams prm, const uint32_t* __restrict__ buf,
                                                                 const uint32_t* __restrict__ flags,
                                                                 const uint32_t* __restrict__ state, int R) {
    __shared__ int s_cnt[32];
    mine_wait(flags, state, R);
    const int v = blockIdx.x, P = prm.P;
    if (threadIdx.x < 32) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const size_t words = static_cast<size_t>(6) * P;
    const uint32_t* base = buf + static_cast<size_t>(state[0] & 1u) * R * words;
    for (int j = threadIdx.x; j < R * P; j += blockDim.x) {
        const int r = j / P, i = j - r * P;
        const int cls = static_cast<int>(base[r * words + static_cast<size_t>(v) * 3 * P + 2 * static_cast<size_t>(P) + i]);
        atomicAdd(&s_cnt[cls & 31], 1);
    }
    __syncthreads();
    if (threadIdx.x < 32) prm.v[v].cls_count[threadIdx.x] = s_cnt[threadIdx.x];
}
// rank_sim / rank_rnd of the LOCAL pixels among all R*P pixels of their class
__global__ void __launch_bounds__(kMineThreads) nce_mine_global_kernel(NceParams prm, const uint32_t* __restrict__ buf,
                                                                       const uint32_t* __restrict__ flags,
                                                                       const uint32_t* __restrict__ state, int R, int rank) {
    const NceView& V = prm.v[blockIdx.z];
    const int P = prm.P, total = R * P;
    __shared__ int s_cls[kMineChunk];
    __shared__ float s_sim[kMineChunk];
    __shared__ uint32_t s_key[kMineChunk];
    mine_wait(flags, state, R);
    const size_t words = static_cast<size_t>(6) * P;
    const uint32_t* base = buf + static_cast<size_t>(state[0] & 1u) * R * words + static_cast<size_t>(blockIdx.z) * 3 * P;
    const int j0 = blockIdx.y * kMineChunk;
    for (int t = threadIdx.x; t < kMineChunk; t += blockDim.x) {
        const int j = j0 + t;
        int cls = -1;
        float sim = 0.f;
        uint32_t key = 0u;
        if (j < total) {
            const int r = j / P, i = j - r * P;
            const uint32_t* src = base + r * words;
            sim = __uint_as_float(src[i]);
            key = src[static_cast<size_t>(P) + i];
            cls = static_cast<int>(src[2 * static_cast<size_t>(P) + i]);
        }
        s_cls[t] = cls; s_sim[t] = sim; s_key[t] = key;
    }
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int cls = V.pl_own[p], gp = rank * P + p;
    const float sim = V.sim[p];
    const uint32_t key = V.rkey[p];
    int rs = 0, rr = 0;
    const int lim = min(kMineChunk, total - j0);
    for (int t = 0; t < lim; ++t) {
        if (s_cls[t] == cls) {
            const int gj = j0 + t;
            rs += (s_sim[t] < sim || (s_sim[t] == sim && gj < gp)) ? 1 : 0;
            rr += (s_key[t] < key || (s_key[t] == key && gj < gp)) ? 1 : 0;
        }
    }
    if (rs) atomicAdd(&V.rank_sim[p], rs);
    if (rr) atomicAdd(&V.rank_rnd[p], rr);
}

// ---- B (large P): the same membership as THRESHOLDS -----------------------------------------------
// rank_rnd < half  <=>  (key, p) < [element at rank half];   kk-half <= rank_sim < kk  <=>  [rank kk-half] <= (sim, p) < [rank kk]
// The three order statistics per class come from the multi-CTA select (segselect.cuh): O(P) per pass instead of O(P^2).
__device__ __forceinline__ unsigned long long mine_key(uint32_t value_bits, int p, int idx_bits) {
    return (static_cast<unsigned long long>(value_bits) << idx_bits) | static_cast<unsigned long long>(p);
}
struct MineSrc {
    const int32_t* pl;
    const float* sim;
    const uint32_t* rkey;
    int P, idx_bits;
    static constexpr int kSel = 3;
    struct Elem { int cls; unsigned long long ks, kr; };
    __device__ long long count() const { return P; }
    __device__ Elem load(long long e) const {
        const int p = static_cast<int>(e);
        return Elem{pl[p], mine_key(ordered_bits(sim[p]), p, idx_bits), mine_key(rkey[p], p, idx_bits)};
    }
    __device__ void get(const Elem& x, long long, int s, int& slot, unsigned long long& key) const {
        slot = s * kNceMaxC + x.cls;
        key = s == 0 ? x.kr : x.ks;
    }
};
// the same over the (sim, key, label) of ALL ranks' pixels as they lie in the exchange buffer (global mining on large loss
// grids): element e = r * P + i is pixel i of rank r, and e is also its index in the canonical (value, global pixel) order
struct MineGlobalSrc {
    const uint32_t* buf;       // this rank's exchange buffer: [2 parities][R][view][sim | key | label][P]
    const uint32_t* state;     // state[0] = step of the current exchange (its parity selects the half)
    size_t words, view_off;    // words per rank (6 P), offset of the view's block (3 P v)
    int P, R, idx_bits;
    static constexpr int kSel = 3;
    struct Elem { int cls; unsigned long long ks, kr; };
    __device__ long long count() const { return static_cast<long long>(R) * P; }
    __device__ Elem load(long long e) const {
        const int r = static_cast<int>(e / P), i = static_cast<int>(e - static_cast<long long>(r) * P);
        const uint32_t* src = buf + (static_cast<size_t>(state[0] & 1u) * R + r) * words + view_off;
        const int ge = static_cast<int>(e);
        return Elem{static_cast<int>(src[2 * static_cast<size_t>(P) + i]),
                    mine_key(ordered_bits(__uint_as_float(src[i])), ge, idx_bits),
                    mine_key(src[static_cast<size_t>(P) + i], ge, idx_bits)};
    }
    __device__ void get(const Elem& x, long long, int s, int& slot, unsigned long long& key) const {
        slot = s * kNceMaxC + x.cls;
        key = s == 0 ? x.kr : x.ks;
    }
};
__global__ void mine_sel_init_kernel(NceParams prm) {       // <<<2, 96>>>
    const NceView& V = prm.v[blockIdx.x];
    const int sel = threadIdx.x >> 5, c = threadIdx.x & 31;
    if (sel >= 3 || c >= kNceMaxC) return;
    const int len = c < prm.C ? V.cls_count[c] : 0;
    const int half = len >> 1;
    const int kk = static_cast<int>(static_cast<double>(len) * 0.6);      // int(len_class * 0.6)
    const int r = sel == 0 ? half : (sel == 1 ? kk - half : kk);
    V.sel_state[sel * kNceMaxC + c] = SegSel{0ull, r, len >= 2 ? 1 : 0};
}

// ---- C: weights, loss partials and gradient w.r.t. the un-normalised features ---------------------
// NPIX = 16: 128 threads (small loss grids, one CTA per 16 pixels keeps all SMs busy at P = 2048).
// NPIX = 64: 512 threads; the prototype rows are read as float4 along d (4 channels per thread and step), which turns the
// shared-memory : FMA instruction ratio from 1:1 into 1:4; same per-element order of operations, so the same values.
template <int NPIX>
__global__ void __launch_bounds__(NPIX * kNceGroups) nce_grad_kernel(NceParams prm) {
    constexpr bool kVec = NPIX > kNcePix;
    extern __shared__ __align__(16) float s_dyn[];
    const NceView& V = prm.v[blockIdx.y];
    const int C = prm.C, D = prm.D, HW = prm.HW, P = prm.P;
    const int pstride = kVec ? D + 4 : D + 1;
    float* s_proto = s_dyn;                                   // [48][pstride]
    float* s_x = s_dyn + kNceSlots * pstride;                 // [D][NPIX] (16-byte aligned: D % 4 == 0)
    float* s_coef = s_x + D * NPIX;                           // [NPIX][49]: d loss / d score per slot (times 1/T)
    __shared__ float s_qdq[NPIX], s_scale[NPIX];
    nce_load_protos(V, s_proto, C, D, pstride);
    __shared__ int s_present, s_len[kNceMaxC];
    __shared__ unsigned long long s_thr[3][kNceMaxC];
    if (threadIdx.x < kNceMaxC) {
        s_len[threadIdx.x] = threadIdx.x < C ? V.cls_count[threadIdx.x] : 0;
        if (prm.idx_bits > 0)
            for (int sel = 0; sel < 3; ++sel) s_thr[sel][threadIdx.x] = V.sel_state[sel * kNceMaxC + threadIdx.x].prefix;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int pres = 0;
        for (int c = 0; c < C; ++c) pres += s_len[c] > 0 ? 1 : 0;
        s_present = pres;
    }
    const int pix = threadIdx.x & (NPIX - 1), grp = threadIdx.x / NPIX;
    // NPIX = 16 is launched with one CTA per tile; the 64-pixel variant is persistent (prototypes staged once per CTA)
    for (int tile = blockIdx.x; tile < prm.nblk_grad; tile += gridDim.x) {
    const int p = tile * NPIX + pix;
    // this thread's per-pixel scalars first, so their loads are in flight together with the x tile's
    float ea[3] = {0.f, 0.f, 0.f}, eb[3] = {0.f, 0.f, 0.f}, so[3] = {0.f, 0.f, 0.f}, st[3] = {0.f, 0.f, 0.f};
    uint32_t mask = 0u;
    int lo = 0, lt = 0;
    float g_sim = 0.f, g_ell = 0.f, g_inv = 0.f;
    uint32_t g_key = 0u;
    if (p < P) {
        lo = V.pl_own[p]; lt = V.pl_oth[p];
        mask = V.setmask[p];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = grp + kNceGroups * j;
            if (c < C) {
                so[j] = V.s_own[static_cast<size_t>(c) * P + p];
                st[j] = V.s_oth[static_cast<size_t>(c) * P + p];
            }
        }
        if (grp == 0) { g_sim = V.sim[p]; g_key = V.rkey[p]; g_ell = V.ell[p]; g_inv = V.inv_norm[p]; }
    }
    if (kVec) nce_tile_load_x<NPIX>(V, s_x, tile * NPIX, D, HW, P);
    else nce_load_x(V, s_x, tile * kNcePix, D, HW, P);
    __syncthreads();
    float intra = 0.f;
    // per-pixel scalars over the pixel's 8 threads (classes grp, grp+8, grp+16), reduced through shared memory in a
    // fixed order
    float* s_part = s_coef + NPIX * 49;                       // [NPIX][kNceGroups][4]
    __shared__ float s_w[NPIX][4];                            // w_x, w_i, lo, lt per pixel
    if (p < P) {
        float z_own = 0.f, z_oth = 0.f, z_set = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = grp + kNceGroups * j;
            if (c < C) {
                ea[j] = expf(so[j] * kInvT);
                eb[j] = expf(st[j] * kInvT);
                z_own += ea[j];
                z_oth += eb[j];
                const int mult = ((mask >> c) & 1u) + (c == lo ? 1 : 0);
                z_set += static_cast<float>(mult) * ea[j];
            }
        }
        float* dstp = s_part + grp * 4 * NPIX + pix;           // [grp][4][NPIX]: conflict-free
        dstp[0] = z_own; dstp[NPIX] = z_oth; dstp[2 * NPIX] = z_set; dstp[3 * NPIX] = 0.f;
    }
    if (grp == 0 && p < P) {
        const int len = s_len[lo];
        const int half = len >> 1;
        int cnt = 0;
        if (len >= 2) {
            if (prm.idx_bits > 0) {
                const unsigned long long ks = mine_key(ordered_bits(g_sim), p + prm.idx_off, prm.idx_bits);
                const unsigned long long kr = mine_key(g_key, p + prm.idx_off, prm.idx_bits);
                cnt = (kr < s_thr[0][lo] ? 1 : 0) + ((ks >= s_thr[1][lo] && ks < s_thr[2][lo]) ? 1 : 0);
            } else {
                const int kk = static_cast<int>(static_cast<double>(len) * 0.6);  // int(len_class * 0.6)
                const int rs = V.rank_sim[p], rr = V.rank_rnd[p];
                cnt = (rr < half ? 1 : 0) + ((rs >= kk - half && rs < kk) ? 1 : 0);
            }
        }
        if (V.mined) V.mined[p] = cnt;
        const float w_x = 0.05f / static_cast<float>(P);                            // 0.1 * mean / 2
        const float w_i = cnt ? prm.intra_scale * 0.05f * static_cast<float>(cnt) /
                                    (static_cast<float>(2 * half) * static_cast<float>(s_present))
                              : 0.f;
        intra = w_i * g_ell;
        s_w[pix][0] = w_x; s_w[pix][1] = w_i;
        s_scale[pix] = g_inv;
    }
    __syncthreads();
    float qdq_part = 0.f;
    if (p < P) {
        float z_own = 0.f, z_oth = 0.f, z_set = 0.f;
#pragma unroll
        for (int g = 0; g < kNceGroups; ++g) {                 // fixed order, identical in all 8 threads
            const float* sp = s_part + g * 4 * NPIX + pix;
            z_own += sp[0]; z_oth += sp[NPIX]; z_set += sp[2 * NPIX];
        }
        const float w_x = s_w[pix][0], w_i = s_w[pix][1];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int c = grp + kNceGroups * j;
            if (c < kSlotHalf) {
                float ac = 0.f, bc = 0.f;
                if (c < C) {
                    const int mult = ((mask >> c) & 1u) + (c == lo ? 1 : 0);
                    ac = w_x * (ea[j] / z_own - (c == lt ? 1.f : 0.f)) +
                         w_i * (static_cast<float>(mult) * ea[j] / z_set - (c == lo ? 1.f : 0.f));
                    bc = w_x * (eb[j] / z_oth - (c == lo ? 1.f : 0.f));
                    qdq_part = fmaf(ac, so[j], qdq_part);
                    qdq_part = fmaf(bc, st[j], qdq_part);
                }
                s_coef[pix * 49 + c] = ac * kInvT;
                s_coef[pix * 49 + kSlotHalf + c] = bc * kInvT;
            }
        }
    }
    __syncthreads();           // everyone has read s_part: reuse it for the qdq partials
    if (p < P) s_part[(grp * 4 + 3) * NPIX + pix] = qdq_part;
    __syncthreads();
    if (grp == 0 && p < P) {
        float qdq = 0.f;
#pragma unroll
        for (int g = 0; g < kNceGroups; ++g) qdq += s_part[(g * 4 + 3) * NPIX + pix];
        s_qdq[pix] = qdq * kInvT;
    }
    __syncthreads();
    if (p < P) {
        const int n = p / HW, q = p - n * HW;
        float coef[kNceSlots];
#pragma unroll
        for (int i = 0; i < kNceSlots; ++i) coef[i] = s_coef[pix * 49 + i];
        const float qdq = s_qdq[pix], inv = s_scale[pix];
        const float gs = prm.grad_scale ? prm.grad_scale[0] : 1.f;
        float* gp = V.grad_x + static_cast<size_t>(n) * D * HW + q;
        // Term order per channel: own c=0, other c=0, own c=1, ...
        if (kVec) {
            // channels 4*dc .. 4*dc+3 with dc = grp + 8*i: a warp is 32 consecutive pixels of one group, so the prototype
            // reads are one broadcast float4 and the stores are 128-byte rows
            for (int dc = grp; dc < (D >> 2); dc += kNceGroups) {
                float dq[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int c = 0; c < kNceMaxC; ++c) {
                    const float4 a = *reinterpret_cast<const float4*>(s_proto + c * pstride + 4 * dc);
                    const float4 b = *reinterpret_cast<const float4*>(s_proto + (kSlotHalf + c) * pstride + 4 * dc);
                    dq[0] = fmaf(coef[c], a.x, dq[0]); dq[0] = fmaf(coef[kSlotHalf + c], b.x, dq[0]);
                    dq[1] = fmaf(coef[c], a.y, dq[1]); dq[1] = fmaf(coef[kSlotHalf + c], b.y, dq[1]);
                    dq[2] = fmaf(coef[c], a.z, dq[2]); dq[2] = fmaf(coef[kSlotHalf + c], b.z, dq[2]);
                    dq[3] = fmaf(coef[c], a.w, dq[3]); dq[3] = fmaf(coef[kSlotHalf + c], b.w, dq[3]);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int d = 4 * dc + u;
                    const float qd = s_x[d * NPIX + pix] * inv;
                    gp[static_cast<size_t>(d) * HW] = (dq[u] - qd * qdq) * inv * gs;
                }
            }
        } else {
            // channel d = grp + 8*i: the 16 pixels of a CTA are contiguous in memory for a fixed d
            for (int d = grp; d < D; d += kNceGroups) {
                float dq = 0.f;
#pragma unroll
                for (int c = 0; c < kNceMaxC; ++c) {
                    dq = fmaf(coef[c], s_proto[c * pstride + d], dq);
                    dq = fmaf(coef[kSlotHalf + c], s_proto[(kSlotHalf + c) * pstride + d], dq);
                }
                const float qd = s_x[d * NPIX + pix] * inv;
                gp[static_cast<size_t>(d) * HW] = (dq - qd * qdq) * inv * gs;
            }
        }
    }
    // fixed-order block partial of the weighted intra loss: the pixel threads (grp 0) are the first NPIX threads
    __shared__ float s_intra[NPIX / 16];
    if (threadIdx.x < NPIX || threadIdx.x < 32) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) intra += __shfl_xor_sync(0xffffffffu, intra, o);
        if ((threadIdx.x & 31) == 0) s_intra[threadIdx.x >> 5] = intra;
    }
    if (NPIX > 32) __syncthreads();
    if (threadIdx.x == 0) {
        float tot = s_intra[0];
        for (int wv = 1; wv < NPIX / 32; ++wv) tot += s_intra[wv];
        V.partial[tile * 4 + 2] = tot;
    }
    if (kVec) __syncthreads();                                 // the next tile rewrites s_x / s_coef / s_part
    }
}

// ---- D: fixed-order final sums (one warp; lane l adds blocks l, l+32, ... then a shuffle tree) -------------
__global__ void __launch_bounds__(192) nce_finalize_kernel(NceParams prm) {
    // warp (v, j) sums partial j of view v: lane l adds blocks l, l+32, ... (four independent loads in flight, added in
    // block order), then a shuffle tree; thread 0 combines
    __shared__ float s_sum[6];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5;
    {
        const int v = wv / 3, j = wv - 3 * v;
        const float* part = prm.v[v].partial + j;
        const int nblk = j == 2 ? prm.nblk_grad : prm.nblk;
        float acc = 0.f;
        int b = lane;
        for (; b + 96 < nblk; b += 128) {
            const float x0 = part[b * 4], x1 = part[(b + 32) * 4], x2 = part[(b + 64) * 4], x3 = part[(b + 96) * 4];
            acc += x0; acc += x1; acc += x2; acc += x3;
        }
        for (; b < nblk; b += 32) acc += part[b * 4];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) s_sum[wv] = acc;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    float s[2][3];
    for (int v = 0; v < 2; ++v)
        for (int j = 0; j < 3; ++j) s[v][j] = s_sum[v * 3 + j];
    const float invP = 1.f / static_cast<float>(prm.P);
    float* L = prm.losses;
    L[0] = s[0][0] * invP;            // nce(q1, proto2, pl1)
    L[1] = s[1][0] * invP;            // nce(q2, proto1, pl2)
    L[2] = s[0][1] * invP;            // nce(q1, proto1, pl2)
    L[3] = s[1][1] * invP;            // nce(q2, proto2, pl1)
    L[4] = s[0][2] / 0.05f;           // loss_intra_nce1 (the partials carry the 0.1/2 weight)
    L[5] = s[1][2] / 0.05f;
    L[6] = 0.1f * (L[0] + L[1]) / 2.f;
    L[7] = 0.1f * (L[2] + L[3]) / 2.f;
    L[8] = 0.1f * (L[4] + L[5]) / 2.f;
    L[9] = L[6] + L[7] + L[8];
}

constexpr int kNceTileMinP = 16384;           // from here on the 64-pixel-tile variants of the pixel / gradient kernels
constexpr int kMinePairwiseMaxP = 16384;      // above this the O(P^2) pairwise ranks lose to the multi-pass select
struct NceLayout {
    size_t per_view, total, clear_bytes;
    size_t off_inv, off_sim, off_ell, off_mask, off_sown, off_soth, off_rs, off_rr, off_cls, off_state, off_hist, off_part;
    int nblk;
};
static NceLayout nce_layout(int N, int C, int h, int w) {
    NceLayout L{};
    const size_t P = static_cast<size_t>(N) * h * w;
    L.nblk = static_cast<int>((P + kNcePix - 1) / kNcePix);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    // what one memset clears: pairwise path = rank arrays + class counts, select path = class counts + state + histograms
    // (both regions are always carved so the size does not depend on the debug switch)
    const size_t small = P <= static_cast<size_t>(kMinePairwiseMaxP) ? P : static_cast<size_t>(kMinePairwiseMaxP);
    L.off_rs = take(small * 4);
    L.off_rr = take(small * 4);
    L.off_cls = take(32 * 4);
    L.off_state = take(3 * kNceMaxC * sizeof(SegSel));
    L.off_hist = take(static_cast<size_t>(3) * kNceMaxC * kSegBins * sizeof(uint32_t));
    L.clear_bytes = off;
    L.off_inv = take(P * 4);
    L.off_sim = take(P * 4);
    L.off_ell = take(P * 4);
    L.off_mask = take(P * 4);
    L.off_sown = take(P * 4 * C);
    L.off_soth = take(P * 4 * C);
    L.off_part = take(static_cast<size_t>(L.nblk) * 4 * 4);
    L.per_view = align_up(off, 1024);
    L.total = 2 * L.per_view;
    return L;
}
size_t contrast_workspace_bytes(int N, int C, int, int h, int w) { return nce_layout(N, C, h, w).total; }

}  // namespace wsegk

using namespace wsegk;

// ---- mining keys on the device -------------------------------------------------------------------
// The "random half" of the hard-pixel mining (contrast_train.py:316,370 draw it with Python's random.sample) only needs one
// uniform 32-bit key per pixel and step; Philox4x32-10 (Salmon et al., SC'11; the Random123 known-answer vectors are checked
// in tests/test_host_logic.py) is counter-based, so key i of stream `stream_id` at `step` is a pure function
//   philox(counter = (i, step_lo, step_hi, stream_id), key = (seed_lo, seed_hi))[0]
// — no generator state, nothing on the host, identical for any sharding of the pixels over ranks (pass the global index).
namespace wsegk {
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
__global__ void __launch_bounds__(256) mining_keys_kernel(uint32_t* __restrict__ out, int count, unsigned long long index0,
                                                          unsigned long long seed, unsigned long long step,
                                                          uint32_t stream_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const unsigned long long gi = index0 + static_cast<unsigned long long>(i);
    // the pixel index takes counter word 0 (and spills into the top half of word 3 beyond 2^32 pixels)
    const uint4 c = make_uint4(static_cast<uint32_t>(gi), static_cast<uint32_t>(step), static_cast<uint32_t>(step >> 32),
                               stream_id ^ (static_cast<uint32_t>(gi >> 32) << 16));
    out[i] = philox4x32_10(c, make_uint2(static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32))).x;
}
}  // namespace wsegk

extern "C" int wsegk_mining_keys(uint32_t* keys, int count, unsigned long long index0, unsigned long long seed,
                                 unsigned long long step, unsigned int stream_id, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(keys != nullptr && count > 0, WSEGK_ERR_BAD_SHAPE, "mining_keys: need a buffer and count > 0");
    mining_keys_kernel<<<ceil_div(count, 256), 256, 0, stream>>>(keys, count, index0, seed, step, stream_id);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

// tests: take the large-plane (multi-CTA select) paths of K3 and K4 whatever the size (wsegk_debug_force_radix)
static thread_local int g_force_radix = 0;   // per host thread: concurrent callers do not see each other's test switches
// tests: run only the first / second half of the global-mining call (wsegk_debug_mine_phases)
static thread_local int g_mine_phases = 3;
extern "C" int wsegk_debug_mine_phases(int mask) {
    if (mask < 1 || mask > 3) {
        set_error("debug_mine_phases: mask must be 1, 2 or 3 (got %d)", mask);
        return WSEGK_ERR_BAD_SHAPE;
    }
    g_mine_phases = mask;
    return WSEGK_OK;
}
extern "C" int wsegk_debug_force_radix(int on) {
    g_force_radix = on ? 1 : 0;
    return WSEGK_OK;
}

static int next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

extern "C" int wsegk_proto_topk_local(const float* norm_cam, const float* f_proj, int pixel_offset, int k,
                                      float* top_val, int32_t* top_idx, float* top_fea, int N, int C, int D, int h,
                                      int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(norm_cam && f_proj && top_val && top_idx && top_fea, WSEGK_ERR_NULL, "proto_topk_local: null pointer");
    WSEGK_REQUIRE(N > 0 && C > 0 && D > 0 && h > 0 && w > 0, WSEGK_ERR_BAD_SHAPE, "proto_topk_local: bad dims");
    const long long P = static_cast<long long>(N) * h * w;
    WSEGK_REQUIRE(k >= 1 && k <= P && k <= kTopkMaxK, WSEGK_ERR_BAD_SHAPE,
                  "proto_topk_local: need 1 <= k <= min(P, %d) (k=%d, P=%lld)", kTopkMaxK, k, P);
    WSEGK_REQUIRE(P + pixel_offset < (1ll << 31), WSEGK_ERR_BAD_SHAPE, "proto_topk_local: pixel index overflow");
    const int kpad = next_pow2(k);
    const size_t smem = static_cast<size_t>(kpad) * sizeof(unsigned long long);
    if (P <= static_cast<long long>(kTopkRegs) * kTopkThreads && !g_force_radix) {
        WSEGK_CUDA_OK(cudaFuncSetAttribute(proto_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        proto_topk_kernel<<<C, kTopkThreads, smem, stream>>>(norm_cam, f_proj, pixel_offset, k, kpad, top_val, top_idx,
                                                             top_fea, N, C, D, h * w);
        count_launches(1);
    } else {
        // multi-CTA select over all N*C*HW values, compaction of the k winners per class, then sort + gather per class
        const TopkLayout L = topk_layout(C);
        WSEGK_REQUIRE(workspace && workspace_bytes >= L.total, WSEGK_ERR_WORKSPACE_TOO_SMALL,
                      "proto_topk_local: workspace %zu < %zu", workspace_bytes, L.total);
        uint8_t* ws = static_cast<uint8_t*>(workspace);
        SegSel* state = reinterpret_cast<SegSel*>(ws + L.off_state);
        int* count = reinterpret_cast<int*>(ws + L.off_count);
        uint32_t* hist = reinterpret_cast<uint32_t*>(ws + L.off_hist);
        unsigned long long* cand = reinterpret_cast<unsigned long long*>(ws + L.off_cand);
        TopkSrc src{norm_cam, C, h * w, index_bits(P), P * C};
        WSEGK_CUDA_OK(cudaMemsetAsync(ws, 0, L.clear_bytes, stream));
        topk_sel_init_kernel<<<1, 32 * ceil_div(C, 32), 0, stream>>>(state, C, k);
        int launches = 1 + seg_select_run(src, src.total, 32 + src.idx_bits, C, state, hist, stream);
        long long blocks = (src.total + 255) / 256;
        if (blocks > 148 * 16) blocks = 148 * 16;
        topk_compact_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(src, state, count, cand, k);
        WSEGK_CUDA_OK(cudaFuncSetAttribute(topk_tail_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        topk_tail_kernel<<<C, kTopkThreads, smem, stream>>>(cand, f_proj, pixel_offset, k, kpad, src.idx_bits, top_val,
                                                            top_idx, top_fea, D, h * w);
        count_launches(launches + 2);
    }
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_proto_finalize(const float* cand_val, const int32_t* cand_idx, const float* cand_fea, int n_cand,
                                    int k, float* proto, int32_t* sel_idx, int C, int D, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(cand_val && cand_idx && cand_fea && proto && sel_idx, WSEGK_ERR_NULL, "proto_finalize: null pointer");
    WSEGK_REQUIRE(C > 0 && D > 0 && D <= 4096 && k >= 1 && k <= n_cand && n_cand <= 16384, WSEGK_ERR_BAD_SHAPE,
                  "proto_finalize: need 1 <= k <= n_cand <= 16384, D <= 4096 (k=%d n_cand=%d D=%d)", k, n_cand, D);
    const int npad = next_pow2(n_cand);
    const size_t smem = static_cast<size_t>(npad) * (sizeof(unsigned long long) + sizeof(uint32_t));
    WSEGK_CUDA_OK(cudaFuncSetAttribute(proto_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    CandSrc src{cand_val, cand_idx, cand_fea, nullptr, n_cand, C, k, D};
    proto_finalize_kernel<<<C, 1024, smem, stream>>>(src, npad, proto, sel_idx, nullptr, nullptr, 0);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

static size_t exchange_cand_floats(int R, int C, int k, int D) {
    return static_cast<size_t>(2) * 2 * R * C * k * (D + 2);
}

extern "C" size_t wsegk_proto_exchange_bytes(int R, int C, int k, int D) {
    if (R <= 0 || C <= 0 || k <= 0 || D <= 0) return 0;
    return align_up(exchange_cand_floats(R, C, k, D) * sizeof(float), 256) + align_up(static_cast<size_t>(2) * R * 4, 256);
}

extern "C" int wsegk_proto_push(const float* top_val, const int32_t* top_idx, const float* top_fea, void* const* peer_bases,
                                int R, int rank, int view, int C, int k, int D, uint32_t* local_state,
                                wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(top_val && top_idx && top_fea && peer_bases && local_state, WSEGK_ERR_NULL, "proto_push: null pointer");
    WSEGK_REQUIRE(R >= 1 && R <= kMaxRanks && rank >= 0 && rank < R && (view == 0 || view == 1) && C > 0 && k > 0 && D > 0 &&
                      (D % 2) == 0,
                  WSEGK_ERR_BAD_SHAPE, "proto_push: need 1 <= R <= %d, 0 <= rank < R, view in {0,1}, even D", kMaxRanks);
    PushPeers peers{};
    const size_t flag_off = align_up(exchange_cand_floats(R, C, k, D) * sizeof(float), 256);
    for (int p = 0; p < R; ++p) {
        WSEGK_REQUIRE(peer_bases[p] != nullptr, WSEGK_ERR_NULL, "proto_push: peer %d has no buffer", p);
        peers.cand[p] = static_cast<float*>(peer_bases[p]);
        peers.flags[p] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(peer_bases[p]) + flag_off);
    }
    proto_push_kernel<<<dim3(kPushChunks, R), 256, 0, stream>>>(top_val, top_idx, top_fea, peers, R, rank, view, C, k, D,
                                                                local_state);
    count_launches(1);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

// optional device buffer of 2 x (1 + 16) uint64 globaltimer stamps written by the flag-wait kernel of each view
static thread_local unsigned long long* g_exchange_stamps = nullptr;
extern "C" int wsegk_debug_exchange_stamps(unsigned long long* device_buffer) {
    g_exchange_stamps = device_buffer;
    return WSEGK_OK;
}

__global__ void stamp_kernel(unsigned long long* dst) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    *dst = t;
}
// diagnostics: *dst = %globaltimer (ns, one clock for all GPUs of a box) when the stream reaches this point
extern "C" int wsegk_debug_stamp(unsigned long long* dst, wsegk_stream_t stream) {
    WSEGK_REQUIRE(dst, WSEGK_ERR_NULL, "debug_stamp: null pointer");
    stamp_kernel<<<1, 1, 0, stream>>>(dst);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_proto_finalize_symm(const void* local_base, int R, int view, int k, float* proto, int32_t* sel_idx,
                                         int C, int D, const uint32_t* local_state, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(local_base && proto && sel_idx && local_state, WSEGK_ERR_NULL, "proto_finalize_symm: null pointer");
    const int n_cand = R * k;
    WSEGK_REQUIRE(R >= 1 && R <= kMaxRanks && (view == 0 || view == 1) && C > 0 && D > 0 && D <= 4096 && k >= 1 &&
                      n_cand <= 16384,
                  WSEGK_ERR_BAD_SHAPE, "proto_finalize_symm: bad dims (R=%d k=%d D=%d)", R, k, D);
    const int npad = next_pow2(n_cand);
    const size_t smem = static_cast<size_t>(npad) * (sizeof(unsigned long long) + sizeof(uint32_t));
    WSEGK_CUDA_OK(cudaFuncSetAttribute(proto_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    const float* base = static_cast<const float*>(local_base);
    const size_t slot_floats = static_cast<size_t>(C) * k * (D + 2);
    const size_t flag_off = align_up(exchange_cand_floats(R, C, k, D) * sizeof(float), 256);
    const uint32_t* flags = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(local_base) + flag_off) + view * R;
    // parity offset is applied inside the kernel (the step number is only known on the device)
    CandSrc src{nullptr, nullptr, nullptr, base + static_cast<size_t>(view) * R * slot_floats, n_cand, C, k, D};
    proto_wait_flags_kernel<<<1, 32, 0, stream>>>(flags, local_state + view * 2, R,
                                                  g_exchange_stamps ? g_exchange_stamps + view * (1 + kMaxRanks) : nullptr);
    proto_finalize_kernel<<<C, 1024, smem, stream>>>(src, npad, proto, sel_idx, flags, local_state + view * 2, R);
    count_launches(2);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

struct GlobalMine {
    void* const* peer_bases;
    const void* local_base;
    int R, rank;
    uint32_t* local_state;
};
static size_t mine_exchange_words(int R, long long P) { return static_cast<size_t>(2) * R * 6 * static_cast<size_t>(P); }

extern "C" size_t wsegk_mine_exchange_bytes(int R, int N, int h, int w) {
    if (R <= 0 || N <= 0 || h <= 0 || w <= 0) return 0;
    return align_up(mine_exchange_words(R, static_cast<long long>(N) * h * w) * 4, 256) + align_up(static_cast<size_t>(R) * 4, 256);
}

static int contrast_loss_impl(const float* x1, const float* x2, const float* proto1, const float* proto2,
                              const int32_t* pl1, const int32_t* pl2, const uint32_t* rand_key1,
                              const uint32_t* rand_key2, float* losses, float* grad_x1, float* grad_x2,
                              int32_t* mined1, int32_t* mined2, const float* grad_scale, int N, int C,
                              int D, int h, int w, void* workspace, size_t workspace_bytes,
                              wsegk_stream_t stream, const GlobalMine* gm) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(x1 && x2 && proto1 && proto2 && pl1 && pl2 && rand_key1 && rand_key2 && losses && grad_x1 &&
                      grad_x2 && workspace,
                  WSEGK_ERR_NULL, "contrast_loss_fwd_bwd: null pointer");
    WSEGK_REQUIRE((D % 32) == 0, WSEGK_ERR_BAD_SHAPE, "contrast_loss_fwd_bwd: D must be a multiple of 32 (got %d)", D);
    WSEGK_REQUIRE(N > 0 && h > 0 && w > 0 && D > 0 && C >= 13 && C <= kNceMaxC, WSEGK_ERR_BAD_SHAPE,
                  "contrast_loss_fwd_bwd: need 13 <= C <= %d (semi-hard mining takes ranks 3..12), got C=%d",
                  kNceMaxC, C);
    const long long P = static_cast<long long>(N) * h * w;
    WSEGK_REQUIRE(P < (1ll << 30), WSEGK_ERR_BAD_SHAPE, "contrast_loss_fwd_bwd: P=%lld pixels is too many", P);
    const bool radix = P > kMinePairwiseMaxP || (g_force_radix && P >= 2);
    if (gm != nullptr) {
        WSEGK_REQUIRE(gm->peer_bases && gm->local_base && gm->local_state && gm->R >= 1 && gm->R <= kMaxRanks &&
                          gm->rank >= 0 && gm->rank < gm->R,
                      WSEGK_ERR_BAD_SHAPE, "contrast_loss_fwd_bwd_global: bad exchange arguments (R=%d rank=%d)", gm->R, gm->rank);
        WSEGK_REQUIRE(P * gm->R < (1ll << 30), WSEGK_ERR_BAD_SHAPE,
                      "contrast_loss_fwd_bwd_global: R*P = %lld pixels is too many", P * gm->R);
    }
    const size_t proto_smem = (static_cast<size_t>(kNceSlots) * (D + 1) + kNcePix * 49 + kNcePix * kNceGroups * 4 +
                               static_cast<size_t>(D) * kNcePix) * sizeof(float);
    WSEGK_REQUIRE(proto_smem <= 200 * 1024, WSEGK_ERR_BAD_SHAPE, "contrast_loss_fwd_bwd: D too large");
    // large grids: 64-pixel tiles (register-tiled dot products, vector shared-memory reads) if they fit
    const size_t tile_pix_smem = (static_cast<size_t>(D) * (kTileProtoStride + kTilePix) + kTilePix * 49 +
                                  kTilePix * kNceGroups * 4 + kTilePix) * sizeof(float);
    const size_t tile_grad_smem = (static_cast<size_t>(kNceSlots) * (D + 4) + static_cast<size_t>(D) * kGradTilePix +
                                   kGradTilePix * 49 + kGradTilePix * kNceGroups * 4) * sizeof(float);
    const bool tiled = (P >= kNceTileMinP || g_force_radix) && tile_pix_smem <= 200 * 1024 && tile_grad_smem <= 200 * 1024;
    const NceLayout L = nce_layout(N, C, h, w);
    WSEGK_REQUIRE(workspace_bytes >= L.total, WSEGK_ERR_WORKSPACE_TOO_SMALL,
                  "contrast_loss_fwd_bwd: workspace %zu < %zu", workspace_bytes, L.total);
    const bool phase_a = gm == nullptr || (g_mine_phases & 1), phase_b = gm == nullptr || (g_mine_phases & 2);
    NceParams prm{};
    prm.N = N; prm.C = C; prm.D = D; prm.HW = h * w; prm.P = static_cast<int>(P);
    prm.nblk = tiled ? static_cast<int>((P + kTilePix - 1) / kTilePix) : L.nblk;
    prm.nblk_grad = tiled ? static_cast<int>((P + kGradTilePix - 1) / kGradTilePix) : L.nblk;
    prm.grad_scale = grad_scale;
    prm.losses = losses;
    prm.idx_bits = radix ? index_bits(gm ? P * gm->R : P) : 0;
    prm.idx_off = (radix && gm) ? gm->rank * static_cast<int>(P) : 0;
    prm.intra_scale = gm ? static_cast<float>(gm->R) : 1.f;
    uint8_t* ws = static_cast<uint8_t*>(workspace);
    for (int v = 0; v < 2; ++v) {
        uint8_t* b = ws + v * L.per_view;
        NceView& V = prm.v[v];
        V.x = v ? x2 : x1;
        V.proto_own = v ? proto2 : proto1;
        V.proto_oth = v ? proto1 : proto2;
        V.pl_own = v ? pl2 : pl1;
        V.pl_oth = v ? pl1 : pl2;
        V.rkey = v ? rand_key2 : rand_key1;
        V.grad_x = v ? grad_x2 : grad_x1;
        V.mined = v ? mined2 : mined1;
        V.inv_norm = reinterpret_cast<float*>(b + L.off_inv);
        V.sim = reinterpret_cast<float*>(b + L.off_sim);
        V.ell = reinterpret_cast<float*>(b + L.off_ell);
        V.setmask = reinterpret_cast<uint32_t*>(b + L.off_mask);
        V.s_own = reinterpret_cast<float*>(b + L.off_sown);
        V.s_oth = reinterpret_cast<float*>(b + L.off_soth);
        V.rank_sim = reinterpret_cast<int*>(b + L.off_rs);
        V.rank_rnd = reinterpret_cast<int*>(b + L.off_rr);
        V.cls_count = reinterpret_cast<int*>(b + L.off_cls);
        V.sel_state = reinterpret_cast<SegSel*>(b + L.off_state);
        V.sel_hist = reinterpret_cast<uint32_t*>(b + L.off_hist);
        V.partial = reinterpret_cast<float*>(b + L.off_part);
        if (!phase_a) continue;
        if (radix)
            WSEGK_CUDA_OK(cudaMemsetAsync(b + L.off_cls, 0, L.clear_bytes - L.off_cls, stream));
        else
            WSEGK_CUDA_OK(cudaMemsetAsync(b + L.off_rs, 0, L.off_state - L.off_rs, stream));
    }
    int grid_a = 0, grid_c = 0;
    if (tiled) {
        WSEGK_CUDA_OK(cudaFuncSetAttribute(nce_pixel_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(tile_pix_smem)));
        WSEGK_CUDA_OK(cudaFuncSetAttribute(nce_grad_kernel<kGradTilePix>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(tile_grad_smem)));
        // persistent CTAs: as many as are co-resident
        int dev = 0, sms = 0, occ_a = 1, occ_c = 1;
        WSEGK_CUDA_OK(cudaGetDevice(&dev));
        WSEGK_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        WSEGK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_a, nce_pixel_tile_kernel, kPixTileThreads, tile_pix_smem));
        WSEGK_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_c, nce_grad_kernel<kGradTilePix>,
                                                                    kGradTilePix * kNceGroups, tile_grad_smem));
        grid_a = std::min(prm.nblk, std::max(1, sms * occ_a));
        grid_c = std::min(prm.nblk_grad, std::max(1, sms * occ_c));
        if (phase_a) nce_pixel_tile_kernel<<<dim3(grid_a, 2), kPixTileThreads, tile_pix_smem, stream>>>(prm);
    } else {
        WSEGK_CUDA_OK(cudaFuncSetAttribute(nce_pixel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(proto_smem)));
        WSEGK_CUDA_OK(cudaFuncSetAttribute(nce_grad_kernel<kNcePix>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(proto_smem)));
        if (phase_a) nce_pixel_kernel<<<dim3(L.nblk, 2), kNceThreads, proto_smem, stream>>>(prm);
    }
    int launches = 3;
    const uint32_t* gbuf = nullptr;
    if (gm != nullptr) {
        // every rank's (sim, key, label) into every peer's buffer; global class sizes
        const int R = gm->R;
        const size_t flag_off = align_up(mine_exchange_words(R, P) * 4, 256);
        MinePeers peers{};
        for (int p = 0; p < R; ++p) {
            WSEGK_REQUIRE(gm->peer_bases[p] != nullptr, WSEGK_ERR_NULL, "contrast_loss_fwd_bwd_global: peer %d has no buffer", p);
            peers.base[p] = static_cast<uint32_t*>(gm->peer_bases[p]);
            peers.flags[p] = reinterpret_cast<uint32_t*>(static_cast<uint8_t*>(gm->peer_bases[p]) + flag_off);
        }
        gbuf = static_cast<const uint32_t*>(gm->local_base);
        const uint32_t* flags = reinterpret_cast<const uint32_t*>(static_cast<const uint8_t*>(gm->local_base) + flag_off);
        if (phase_a) mine_push_kernel<<<dim3(kPushChunks, R), 256, 0, stream>>>(prm, peers, R, gm->rank, gm->local_state);
        if (!phase_b) {
            count_launches(2);
            WSEGK_CUDA_OK(cudaGetLastError());
            return WSEGK_OK;
        }
        mine_global_count_kernel<<<2, 1024, 0, stream>>>(prm, gbuf, flags, gm->local_state, R);   // waits for the peers' flags
        launches += 2;
        if (!radix) {
            nce_mine_global_kernel<<<dim3(ceil_div(prm.P, kMineThreads), ceil_div(R * prm.P, kMineChunk), 2), kMineThreads, 0,
                                     stream>>>(prm, gbuf, flags, gm->local_state, R, gm->rank);
            ++launches;
        }
    }
    if (radix) {
        mine_sel_init_kernel<<<2, 96, 0, stream>>>(prm);
        ++launches;
        for (int v = 0; v < 2; ++v) {
            const NceView& V = prm.v[v];
            if (gm != nullptr) {
                // thresholds over ALL ranks' pixels (every rank derives the same three keys per class); the parity half of
                // the buffer is a device-side value (step counter), so the source takes it from there
                const MineGlobalSrc src{gbuf, gm->local_state, static_cast<size_t>(6) * prm.P, static_cast<size_t>(3) * prm.P * v,
                                        prm.P, gm->R, prm.idx_bits};
                launches += seg_select_run(src, static_cast<long long>(gm->R) * prm.P, 32 + prm.idx_bits, 3 * kNceMaxC,
                                           V.sel_state, V.sel_hist, stream);
            } else {
                const MineSrc src{V.pl_own, V.sim, V.rkey, prm.P, prm.idx_bits};
                launches += seg_select_run(src, prm.P, 32 + prm.idx_bits, 3 * kNceMaxC, V.sel_state, V.sel_hist, stream);
            }
        }
    } else if (gm == nullptr) {
        nce_mine_kernel<<<dim3(ceil_div(prm.P, kMineThreads), ceil_div(prm.P, kMineChunk), 2), kMineThreads, 0,
                          stream>>>(prm);
        ++launches;
    }
    if (tiled)
        nce_grad_kernel<kGradTilePix><<<dim3(grid_c, 2), kGradTilePix * kNceGroups, tile_grad_smem, stream>>>(prm);
    else
        nce_grad_kernel<kNcePix><<<dim3(L.nblk, 2), kNceThreads, proto_smem, stream>>>(prm);
    nce_finalize_kernel<<<1, 32 * 6, 0, stream>>>(prm);
    count_launches(launches);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_contrast_loss_fwd_bwd(const float* x1, const float* x2, const float* proto1, const float* proto2,
                                           const int32_t* pl1, const int32_t* pl2, const uint32_t* rand_key1,
                                           const uint32_t* rand_key2, float* losses, float* grad_x1, float* grad_x2,
                                           int32_t* mined1, int32_t* mined2, const float* grad_scale, int N, int C,
                                           int D, int h, int w, void* workspace, size_t workspace_bytes,
                                           wsegk_stream_t stream) {
    return contrast_loss_impl(x1, x2, proto1, proto2, pl1, pl2, rand_key1, rand_key2, losses, grad_x1, grad_x2, mined1, mined2,
                              grad_scale, N, C, D, h, w, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int wsegk_contrast_loss_fwd_bwd_global(const float* x1, const float* x2, const float* proto1, const float* proto2,
                                                  const int32_t* pl1, const int32_t* pl2, const uint32_t* rand_key1,
                                                  const uint32_t* rand_key2, float* losses, float* grad_x1, float* grad_x2,
                                                  int32_t* mined1, int32_t* mined2, const float* grad_scale, int N, int C,
                                                  int D, int h, int w, void* workspace, size_t workspace_bytes,
                                                  void* const* peer_bases, const void* local_base, int R, int rank,
                                                  uint32_t* local_state, wsegk_stream_t stream) {
    const GlobalMine gm{peer_bases, local_base, R, rank, local_state};
    return contrast_loss_impl(x1, x2, proto1, proto2, pl1, pl2, rand_key1, rand_key2, losses, grad_x1, grad_x2, mined1, mined2,
                              grad_scale, N, C, D, h, w, workspace, workspace_bytes, stream, &gm);
}
