// Multi-CTA "element at rank r" over many independent sets at once (radix select by value, 11-bit digits, MSD first).
//
// Used where one CTA per set does not scale: the per-class top-k of K3 (contrast_train.py:202,231) and the per-class
// order statistics of the hard-pixel mining in K4 (:302-322) once the loss grid has more than a few thousand pixels.
// Every element carries a composite key (value bits above its flat index), so keys are unique and the element at a
// rank is well defined: the canonical (value, lowest index) order of the rest of the library.
//
// One pass = seg_hist_kernel (all elements, warp-aggregated integer atomics into [slot][2048] histograms) followed by
// seg_scan_kernel (one CTA per slot: find the digit holding the wanted rank, extend the prefix, clear the histogram).
// After the last pass SegSel::prefix IS the key at the wanted rank.  Integer arithmetic only: bit-reproducible.
#pragma once
#include "common.cuh"

namespace wsegk {

constexpr int kSegBits = 11;
constexpr int kSegBins = 1 << kSegBits;
constexpr int kSegThreads = 256;
constexpr int kSegMaxPasses = 6;      // keys of up to 63 bits

struct SegSel {
    unsigned long long prefix;        // the key bits decided so far (after the last pass: the key at the wanted rank)
    int k_rem;                        // 0-based ascending rank wanted inside the bucket `prefix`
    int live;                         // 0: nothing to select for this slot
};

struct SegPasses {
    int n;
    int shift[kSegMaxPasses], width[kSegMaxPasses];
};
inline SegPasses seg_passes(int total_bits) {
    SegPasses p{};
    int hi = total_bits;
    while (hi > 0 && p.n < kSegMaxPasses) {
        const int sh = hi > kSegBits ? hi - kSegBits : 0;
        p.shift[p.n] = sh;
        p.width[p.n] = hi - sh;
        ++p.n;
        hi = sh;
    }
    return p;
}
// bits needed for indices 0 .. count-1
inline int index_bits(long long count) {
    int b = 1;
    while ((1ll << b) < count) ++b;
    return b;
}

// Src: struct with  static constexpr int kSel;  struct Elem;  __device__ long long count() const;
//                   __device__ Elem load(long long e) const;
//                   __device__ void get(const Elem&, long long e, int s, int& slot, unsigned long long& key) const;
template <class Src>
__global__ void __launch_bounds__(kSegThreads) seg_hist_kernel(Src src, const SegSel* __restrict__ state,
                                                               uint32_t* __restrict__ hist, int shift, int width) {
    const long long count = src.count();
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const int lane = threadIdx.x & 31;
    // the loop bound is warp-uniform so the full-mask ballots below are safe
    for (long long base = static_cast<long long>(blockIdx.x) * blockDim.x + (threadIdx.x - lane); base < count;
         base += stride) {
        const long long e = base + lane;
        const bool in = e < count;
        typename Src::Elem el{};
        if (in) el = src.load(e);
#pragma unroll
        for (int s = 0; s < Src::kSel; ++s) {
            int slot = -1;
            unsigned long long key = 0ull;
            if (in) src.get(el, e, s, slot, key);
            bool act = false;
            uint32_t addr = 0u;
            if (slot >= 0) {
                const SegSel st = state[slot];
                act = st.live != 0 && (key >> (shift + width)) == st.prefix;
                addr = static_cast<uint32_t>(slot) * kSegBins +
                       static_cast<uint32_t>((key >> shift) & ((1ull << width) - 1ull));
            }
            const unsigned am = __ballot_sync(0xffffffffu, act);
            if (act) {
                const unsigned peers = __match_any_sync(am, addr);
                if (lane == __ffs(peers) - 1) atomicAdd(hist + addr, static_cast<uint32_t>(__popc(peers)));
            }
        }
    }
}

// one CTA per slot; thread t owns bins 8t .. 8t+7
static __global__ void __launch_bounds__(kSegThreads) seg_scan_kernel(SegSel* __restrict__ state, uint32_t* __restrict__ hist,
                                                               int width) {
    __shared__ int s_warp[kSegThreads / 32];
    const int slot = blockIdx.x;
    const SegSel st = state[slot];
    if (!st.live) return;                                   // uniform: nothing was counted for this slot
    uint4* h = reinterpret_cast<uint4*>(hist + static_cast<size_t>(slot) * kSegBins);
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint4 a = h[2 * t], b = h[2 * t + 1];
    h[2 * t] = make_uint4(0u, 0u, 0u, 0u);                  // leave the histogram clean for the next pass
    h[2 * t + 1] = make_uint4(0u, 0u, 0u, 0u);
    const int local[8] = {static_cast<int>(a.x), static_cast<int>(a.y), static_cast<int>(a.z), static_cast<int>(a.w),
                          static_cast<int>(b.x), static_cast<int>(b.y), static_cast<int>(b.z), static_cast<int>(b.w)};
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += local[j];
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    int excl = incl - sum;
    for (int wv = 0; wv < warp; ++wv) excl += s_warp[wv];
    const int need = st.k_rem;
    if (excl <= need && need < excl + sum) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (need < cum + local[j]) {
                state[slot].prefix = (st.prefix << width) | static_cast<unsigned long long>(8 * t + j);
                state[slot].k_rem = need - cum;
                break;
            }
            cum += local[j];
        }
    }
}

// all passes of one selection problem, on `stream`; returns the number of kernels launched
template <class Src>
inline int seg_select_run(const Src& src, long long count, int total_bits, int n_slots, SegSel* state, uint32_t* hist,
                          cudaStream_t stream) {
    const SegPasses ps = seg_passes(total_bits);
    long long blocks = (count + kSegThreads - 1) / kSegThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    for (int j = 0; j < ps.n; ++j) {
        seg_hist_kernel<Src><<<static_cast<unsigned>(blocks), kSegThreads, 0, stream>>>(src, state, hist, ps.shift[j],
                                                                                        ps.width[j]);
        seg_scan_kernel<<<n_slots, kSegThreads, 0, stream>>>(state, hist, ps.width[j]);
    }
    return 2 * ps.n;
}

}  // namespace wsegk
