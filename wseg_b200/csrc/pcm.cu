// K2 — SEAM pixel correlation module (replaces network/resnet38_contrast.py:70-73 after the f9 conv).
//
//   fh_i = f_i / (|f_i| + 1e-5)              (per pixel, over K channels)
//   S    = fh fh^T  (HW x HW, symmetric),  A = relu(S),  D_j = sum_i A_ij + 1e-5
//   out[c,j] = sum_i cam[c,i] A_ij / D_j
//
// Design (B200).  The HW x HW affinity never exists in HBM.  A CTA PAIR (thread-block cluster of 2, tcgen05
// cta_group::2, UMMA M = 256) owns 256 "j" pixels and sweeps the "i" pixels in stages of 2*wb pixels: each CTA
// bulk-copies ONE wb-row tile per stage (half the L2 traffic of two independent CTAs), the leader CTA issues
//   S^T tile [256 x 2wb] = Fj Fi^T                      (SS MMA, FP16 operands, FP32 accumulators in TMEM)
// four warps per CTA turn their 128 TMEM lanes into relu(S) as packed FP16 IN PLACE (aliasing the accumulator)
// and record the sign bits of S, and a second MMA with the A operand in TMEM multiplies by the
// [cam_hi ; 1 | cam_lo] rows of the same pixels, so numerators and the column sum D_j share one accumulator.
//
// Precision modes (wsegk_pcm_set_mode).  "precise" (default) splits fh into FP16 hi + residual lo and accumulates
// hi*hi + hi*lo + lo*hi (about 2^-21 relative), because the BACKWARD needs the sign of S: d relu is a step, and a
// sign flipped by operand rounding adds or drops a whole term of the gradient.  The forward stores the sign bits
// (16 bytes per row and stage); the backward reads them instead of recomputing S, so it only does the rank-(2C+2)
// product T = U V^T and dFh_j += (mask .* T) Fi on the tensor cores.  "fast" uses single FP16 products.
//
// Operands are stored by the prep kernels directly in the UMMA canonical no-swizzle layout, so every tile is one
// contiguous bulk copy (cp.async.bulk, no tensor maps).
//
// Schedule.  Both main kernels are persistent: 74 CTA pairs, each running the stage list PcmSched (common.cuh) gives it —
// whole (image, j pair) units while they divide evenly ("rows": all pairs on consecutive units, so the i-side tiles are
// shared through L2), then an equal contiguous share of the remaining units' stages (stream-K tail).  A share that crosses
// a unit boundary starts a new segment: new A-side tile (after the a_free / u_free commit of the old segment's last
// product), accumulator restart, own partial slot.
#include <algorithm>
#include <cstdlib>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace wsegk {

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
// Canonical K-major tile of R rows x KK 16-bit columns: 16-byte chunk (row r, column group g) lives at
// chunk index g*R + r, i.e. core matrices (8 rows x 16 B) are 128 contiguous bytes, SBO = 128 B, LBO = R*16 B.

__device__ __forceinline__ uint32_t pack_h2(__half a, __half b) {
    return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

// f [N,K,HW] fp32 -> normalised features as A tiles (128 rows) and B tiles (wb rows) in FP16, + inv_norm.
// One CTA per (128 pixels, image), 256 threads: thread (pixel, half) keeps its K/2 channel values in registers
// (channel groups half, half+2, ...), so global loads are coalesced along the pixel axis, the 16-byte chunks a thread
// writes are already in tile order (row = pixel), and no shared-memory transpose is needed.
// kPrecise: also the fp32 values pixel-major (fh32 [n][p][K], what pcm_sign_fixup_kernel re-evaluates S from) and the
// batch maximum of |fh_p - fp16(fh_p)|^2 (lo2max, atomicMax on the float bits) that sets the fix-up threshold.
// channel parts per pixel of the feature preparation = CTA size / 128.  Two parts: 256 threads, 96 values per thread at
// K = 192, ~3 CTAs of 8 warps per SM.  -DWSEGK_PREP_PARTS=4 (512 threads, 48 values per thread) was measured and is no
// faster (view-1 forward chain 107.4 vs 109.2 us, step 0.301 vs 0.309 ms): the kernel is not short of warps.
#ifndef WSEGK_PREP_PARTS
#define WSEGK_PREP_PARTS 2
#endif
constexpr int kPrepParts = WSEGK_PREP_PARTS;
constexpr int kPrepThreads = 128 * kPrepParts;
template <bool kPrecise>
__device__ __forceinline__ void pcm_prep_f_body(const float* __restrict__ f, uint4* __restrict__ fa, uint4* __restrict__ fb,
                                                float* __restrict__ fh32, unsigned int* __restrict__ lo2max,
                                                float* __restrict__ inv_norm, int K, int HW, int PJ, int NS, int wb,
                                                int t, int n) {
    constexpr int GMAX = kMaxPcmK / (8 * kPrepParts);          // channel groups per thread
    __shared__ float s_part[kPrepParts][kTile];
    __shared__ float s_lo2[kPrepParts][kTile];
    const int px = threadIdx.x & 127, half = threadIdx.x >> 7;  // "half": which of the kPrepParts interleaved sets of channel groups
    const int p = t * kTile + px;
    const int HWpj = PJ * 2 * kTile, HWpi = NS * 2 * wb;
    const int groups = K >> 3;
    const float* src = f + static_cast<size_t>(n) * K * HW + p;
    float v[GMAX][8];
    float ss = 0.f;
#pragma unroll
    for (int gi = 0; gi < GMAX; ++gi) {
        const int g = half + kPrepParts * gi;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            v[gi][e] = (g < groups && p < HW) ? __ldg(src + static_cast<size_t>(g * 8 + e) * HW) : 0.f;
            ss = fmaf(v[gi][e], v[gi][e], ss);
        }
    }
    s_part[half][px] = ss;
    __syncthreads();
    float ssum = 0.f;
#pragma unroll
    for (int q = 0; q < kPrepParts; ++q) ssum += s_part[q][px];        // fixed order
    const float nrm = sqrtf(ssum);
    const float inv = 1.f / (nrm + 1e-5f);
    if (half == 0 && p < HWpj) inv_norm[static_cast<size_t>(n) * HWpj + p] = (p < HW) ? inv : 0.f;
    const bool in_a = p < HWpj, in_b = p < HWpi;
    const int tb = p / wb, rb = p - tb * wb;
    const size_t base_a = (static_cast<size_t>(n) * 2 * PJ + t) * groups * kTile + px;
    const size_t base_b = (static_cast<size_t>(n) * 2 * NS + tb) * groups * wb + rb;
    float4* dst32 = reinterpret_cast<float4*>(fh32 + (static_cast<size_t>(n) * HWpj + p) * K);
    float lo2 = 0.f;
#pragma unroll
    for (int gi = 0; gi < GMAX; ++gi) {
        const int g = half + kPrepParts * gi;
        if (g >= groups) break;
        __half hi[8];
        float x[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            x[e] = v[gi][e] * inv;
            hi[e] = __float2half_rn(x[e]);
            const float lo = x[e] - __half2float(hi[e]);
            lo2 = fmaf(lo, lo, lo2);
        }
        const uint4 ohi = make_uint4(pack_h2(hi[0], hi[1]), pack_h2(hi[2], hi[3]), pack_h2(hi[4], hi[5]),
                                     pack_h2(hi[6], hi[7]));
        if (in_a) {
            fa[base_a + static_cast<size_t>(g) * kTile] = ohi;
            if (kPrecise) {
                dst32[2 * g] = make_float4(x[0], x[1], x[2], x[3]);
                dst32[2 * g + 1] = make_float4(x[4], x[5], x[6], x[7]);
            }
        }
        if (in_b) fb[base_b + static_cast<size_t>(g) * wb] = ohi;
    }
    if (kPrecise) {
        s_lo2[half][px] = lo2;
        __syncthreads();
        if (half == 0) {
            float m = 0.f;
#pragma unroll
            for (int q = 0; q < kPrepParts; ++q) m += s_lo2[q][px];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
            if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(lo2max, __float_as_uint(m));   // non-negative floats order like their bits
        }
    }
}

// cam_norm [N,C,HW] fp32 -> per (image, stage) two operand blocks of [32 rows x 2wb pixels] fp16, K-major (K = pixel):
//   rank 0 block: rows 0..C-1 = fp16(cam), row C = 1 (column-sum row), others 0;   rank 1 block: rows 0..C-1 = cam - hi.
// kLogits: `cam` holds the RAW class logits and `peak` the per-plane maxima (cam_plane_peak_kernel); the in-network
// normalisation (K1a, network/resnet38_contrast.py:41-48) is applied here, with cam_pixel_kernel's arithmetic, and also
// written to cam_norm_out — K1a's per-pixel pass leaves the critical path in front of the PCM.
template <bool kLogits>
__device__ __forceinline__ void pcm_prep_cam_body(const float* __restrict__ cam, const float* __restrict__ peak,
                                                  float* __restrict__ cam_norm_out, uint4* __restrict__ camb,
                                                  int C, int HW, int NS, int wb, int s, int n) {
    __shared__ __align__(16) uint16_t s_blk[2 * 16 * kCamRows * 8];   // [rank][pixel group g][row][8]
    __shared__ float s_peak[kCamRows];
    const int q = threadIdx.x;
    const int ncol = 2 * wb;
    const int p = s * ncol + q;
    const bool live = q < ncol && p < HW;
    const float* src = cam + static_cast<size_t>(n) * C * HW + p;
    if (kLogits) {
        if (threadIdx.x < kCamRows) s_peak[threadIdx.x] = threadIdx.x < C ? peak[n * C + threadIdx.x] : 1.f;
        __syncthreads();
    }
    if (q < ncol) {
        const int g = q >> 3, e = q & 7;
        const int G = ncol >> 3;
        float y[kCamRows];
#pragma unroll
        for (int c0 = 0; c0 < kCamRows; c0 += 8) {            // eight independent loads in flight
#pragma unroll
            for (int j = 0; j < 8; ++j) y[c0 + j] = (live && c0 + j < C) ? __ldg(src + static_cast<size_t>(c0 + j) * HW) : 0.f;
        }
        if (kLogits) {
            float fg = 0.f;
#pragma unroll
            for (int c = 1; c < kCamRows; ++c) {
                y[c] = c < C ? fmaxf(fmaxf(y[c], 0.f) - 1e-5f, 0.f) / s_peak[c] : 0.f;
                fg = fmaxf(fg, y[c]);
            }
            y[0] = 1.f - fg;
#pragma unroll
            for (int c = 1; c < kCamRows; ++c) y[c] = (y[c] < fg) ? 0.f : y[c];
            if (live) {
                float* dstn = cam_norm_out + static_cast<size_t>(n) * C * HW + p;
#pragma unroll
                for (int c = 0; c < kCamRows; ++c)
                    if (c < C) dstn[static_cast<size_t>(c) * HW] = y[c];
            }
        }
#pragma unroll
        for (int c = 0; c < kCamRows; ++c) {
            float hi = 0.f, lo = 0.f;
            if (live && c < C) {
                hi = __half2float(__float2half_rn(y[c]));
                lo = y[c] - hi;
            } else if (live && c == C) {
                hi = 1.f;
            }
            s_blk[((0 * G + g) * kCamRows + c) * 8 + e] = __half_as_ushort(__float2half_rn(hi));
            s_blk[((1 * G + g) * kCamRows + c) * 8 + e] = __half_as_ushort(__float2half_rn(lo));
        }
    }
    __syncthreads();
    const int chunks = 2 * (ncol >> 3) * kCamRows;                 // uint4 chunks in both blocks
    uint4* dst = camb + (static_cast<size_t>(n) * NS + s) * chunks;
    const uint4* s4 = reinterpret_cast<const uint4*>(s_blk);
    for (int i = threadIdx.x; i < chunks; i += blockDim.x) dst[i] = s4[i];
}

// both operand preparations in ONE launch (they are independent; as two launches the 7 us cam preparation sat on the
// critical path behind the feature preparation): blocks [0, tiles) prepare features, blocks [tiles, tiles + NS) cam blocks
template <bool kPrecise, bool kLogits>
__global__ void __launch_bounds__(kPrepThreads) pcm_prep_kernel(const float* __restrict__ f, uint4* __restrict__ fa,
                                                       uint4* __restrict__ fb, float* __restrict__ fh32,
                                                       unsigned int* __restrict__ lo2max, float* __restrict__ inv_norm,
                                                       const float* __restrict__ cam, const float* __restrict__ peak,
                                                       float* __restrict__ cam_norm_out, uint4* __restrict__ camb, int C,
                                                       int K, int HW, int PJ, int NS, int wb, int tiles) {
    if (static_cast<int>(blockIdx.x) < tiles)
        pcm_prep_f_body<kPrecise>(f, fa, fb, fh32, lo2max, inv_norm, K, HW, PJ, NS, wb, blockIdx.x, blockIdx.y);
    else
        pcm_prep_cam_body<kLogits>(cam, peak, cam_norm_out, camb, C, HW, NS, wb, static_cast<int>(blockIdx.x) - tiles,
                                   blockIdx.y);
}

// ------------------------------------------------------------------------------------------------
// forward main kernel (CTA pair)
// ------------------------------------------------------------------------------------------------
// 12 warps.  warps 0-7: TMEM warps — two per TMEM lane quarter; warp 8: bulk-copy producer of the feature tiles, warp 9:
// issuer of the S products (leader CTA) / relay (peer CTA), warp 10: bulk-copy producer of the cam blocks, warp 11: issuer
// of the P V products (leader CTA)
// (a warp may only touch lanes 32*(warp%4) ..), the first converts S columns [0,64) of its 32 rows, the second the rest.
// One warp keeps only ~2 bulk copies in flight (tools/ubench/bulkcp2.cu: 630 clk per copy per issuing warp whatever its
// size, perfectly overlapped across warps), so every stream of tiles gets its own issuing warp.
//
// MODE 0 (inference): P = relu(S) only.  MODE 1 ("fast"): + the sign bits of S for the backward.  MODE 2 ("precise"):
// + a second bit per entry, |S| < tau, where tau bounds what the FP16 rounding of the operands can do to S; those entries
// (~0.2 %) are re-evaluated in fp32 by pcm_sign_fixup_kernel, which patches the sign mask.  Round 1 got the same signs from
// three MMA chains (hi*hi + hi*lo + lo*hi): 3x the tensor work for a correction that matters for 1e-4 of the entries.
// backward main kernel: 4 TW TMEM warps (TW per TMEM lane quarter), then the U/V producer, the issuer (relay in the peer CTA) and
// the two producers of the channel halves of F.  TW = 2 (default): each warp converts half of a stage's columns.  TW = 4
// (-DWSEGK_BWD_TW=4, 640 threads, 96 registers): a quarter each.  Measured round 3 at N = 8, 56x56: the SAME 63-65 us — a TMEM
// warp's stage costs ~1300 clk whether it owns 64 or 32 columns (mbarrier poll, tcgen05.ld, tcgen05.st + wait, arrive:
// fixed latencies), and with two T / E buffers the ring T(i) -> convert -> G(i) -> T(i + 2) costs the same ~3000 clk per two
// stages.  A third buffer alone (48-row tiles) changed nothing either: both limits sit at ~1470 clk per stage against 1283
// of tensor-pipe work, so only three buffers AND two stages in flight in the TMEM warps (as in the forward) would
// reach the pipe, and three buffers need 96-column stages, which cost the forward 4 us.  Left as it is.
#ifndef WSEGK_BWD_TW
#define WSEGK_BWD_TW 2
#endif
constexpr int kBwdTW = WSEGK_BWD_TW;
constexpr int kBwdTmemWarps = 4 * kBwdTW;
constexpr int kMainThreads = (kBwdTmemWarps + 4) * 32;
constexpr int kBwdWarpProdUV = kBwdTmemWarps, kBwdWarpIssue = kBwdTmemWarps + 1, kBwdWarpProdX = kBwdTmemWarps + 2;
constexpr int kFwdThreads = 384;
// forward kernel roles.  The TMEM warps come FIRST: the sub-partition arbiter favours the higher warp ids, and the two ALU-dense
// TMEM warps of a sub-partition otherwise starve the producer / issuer warp that shares it (trace r2: the S issuer needed
// 380 clk after its product had completed to get to its next wait)
constexpr int kFwdTmemWarp0 = 0;       // TMEM warps 0..7
constexpr int kFwdWarpProdF = 8, kFwdWarpS = 9, kFwdWarpProdC = 10, kFwdWarpPV = 11;
constexpr int kFStages = 4;
constexpr int kCamStages = 4;
constexpr uint32_t kTmemCols = 512;
// forward: S / P buffers in TMEM.  The hand-off chain S(i) -> convert -> "P ready" arrive -> P V(i) -> commit -> S(i + SB) takes
// ~4200 clk (trace r2/r3: the remote arrive alone ~1400 behind the warp's mask stores) against ~1020 clk of tensor-pipe work
// per stage, so the ring needs FOUR buffers to keep the pipe fed; four fit beside the 64 accumulator columns when a buffer
// is exactly its 2*WB columns wide (4 x 112 + 64 = 512 for 56-row tiles, 4 x 96 + 64 for 48), three of 128 otherwise.
constexpr int kSBufMax = 4;
// forward TMEM warps: true = the two warps of a lane quarter take ALTERNATE STAGES (each converts all 2*WB columns of its
// stage), false = both work on every stage, splitting its columns (round 2).  A stage costs a TMEM warp ~1300 clk of mostly
// latency (barrier poll, tcgen05.ld, the dependent convert chain, tcgen05.st + wait, the arrive) against ~1020 clk of
// tensor-pipe work, and with all eight warps on the same stage those latencies are serial; with two stages in flight per
// sub-partition they overlap.
#ifdef WSEGK_FWD_NOPP
constexpr bool kFwdPingPong = false;
#else
constexpr bool kFwdPingPong = true;
#endif
#ifdef WSEGK_FWD_A1
constexpr int kABuf = 1;            // ablation: one A-side tile buffer
#else
constexpr int kABuf = 2;            // A-side (j) tile buffers: segment k uses slot k % 2
#endif
template <int WB>
struct FwdTmem {
#ifdef WSEGK_FWD_SBUF3
    static constexpr int SB = 3;
#else
    static constexpr int SB = (2 * WB <= 112) ? 4 : 3;
#endif
    static constexpr uint32_t STRIDE = SB == 4 ? 2 * WB : 128;     // TMEM columns between consecutive S / P buffers
    static constexpr uint32_t COL_O = SB * STRIDE;                  // 64 fp32 columns [num_hi | colsum | num_lo]
    static_assert(COL_O + 64 <= 512, "TMEM columns");
};
constexpr uint32_t kColG = 256;     // backward: K (<= 192) fp32 columns dFh accumulator
constexpr int kSplitCol = 64;       // S columns [0,64) belong to the first TMEM warp of a quarter, [64, NCOL) to the second

struct FwdBars {
    // full_* / fj of the LEADER also count one arrival relayed from the peer CTA ("my copy of this stage landed too")
    uint64_t fj[2], a_free[2], full_f[kFStages], empty_f[kFStages], full_c[kCamStages], empty_c[kCamStages], s_full[kSBufMax],
        p_full[kSBufMax], o_full, o_free;
    uint32_t tmem_base;
};

__device__ __forceinline__ __half2 as_half2(uint32_t u) { return *reinterpret_cast<__half2*>(&u); }

__device__ __forceinline__ uint32_t hfma2_relu_u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hmin2_u(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
// 1.0 / 0.0 per half: a > |b|
__device__ __forceinline__ uint32_t hset_gt_abs_u(uint32_t a, uint32_t b) {
    uint32_t d, t;
    asm("abs.f16x2 %0, %1;" : "=r"(t) : "r"(b));              // ptxas folds it into the compare's operand modifier
    asm("set.gt.f16x2.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(t));
    return d;
}

// NV fp32 accumulator values -> relu as packed fp16 (pk), and per value one bit in `sbits` ([S > 0]) and one in `nbits`
// ([|S| < tau]): bit m belongs to value 2m, bit 16+m to value 2m+1.  Both tests are made on the fp16-rounded value h:
// "S rounds to a positive fp16" — values below 2^-25 count as zero, far inside the accumulation error of S.
// The two TMEM warps of a sub-partition are bound by its ALU pipe (one warp instruction per 2 clk; tools/ubench/alu_rates.cu:
// the first sequence — cvt, 2 compares to masks, 3 LOP3 — cost 12 clk per pair), so the work is spread over BOTH pipes,
// 3 + 3 instructions per pair (6-7 clk):
//   ALU pipe: cvt h;  ind = min(p, 2^-24) = 0x0001 per positive half;  near = set(tau > |h|) = 1.0 / 0.0 per half
//   FMA pipe: p = fma.relu(h, 1, 0);  sbits += ind << m (IMAD);  nacc += near * 2^(m mod 8) (HFMA2, exact fp16 integers < 256)
template <int NV, int MODE>
__device__ __forceinline__ void relu_pack_bits(const uint32_t* r, uint32_t* pk, uint32_t tau2, uint32_t& sbits,
                                               uint32_t& nbits) {
    sbits = 0;
    nbits = 0;
    uint32_t nacc[2] = {0u, 0u};
#pragma unroll
    for (int m = 0; m < NV / 2; ++m) {
        if (MODE == 0) {
            pk[m] = pack_f16x2_relu(__uint_as_float(r[2 * m]), __uint_as_float(r[2 * m + 1]));
        } else if (MODE == 1) {
            const uint32_t p = pack_f16x2_relu(__uint_as_float(r[2 * m]), __uint_as_float(r[2 * m + 1]));
            pk[m] = p;
            sbits = hmin2_u(p, 0x00010001u) * (1u << m) + sbits;
        } else {
            const uint32_t h = pack_f16x2(__uint_as_float(r[2 * m]), __uint_as_float(r[2 * m + 1]));
            const uint32_t p = hfma2_relu_u(h, 0x3c003c00u, 0u);
            pk[m] = p;
            sbits = hmin2_u(p, 0x00010001u) * (1u << m) + sbits;
            const uint32_t w = (0x3c00u + (static_cast<uint32_t>(m & 7) << 10)) * 0x00010001u;      // 2^(m mod 8) in both halves
            nacc[m >> 3] = hfma2_u(hset_gt_abs_u(tau2, h), w, nacc[m >> 3]);
        }
    }
    if (MODE == 2) {
        // fp16 integers 0..255 -> bits: adding 1024 leaves the integer in the low mantissa bits
        const uint32_t k1024 = 0x64006400u, one = 0x3c003c00u;
        nbits = hfma2_u(nacc[0], one, k1024) & 0x00ff00ffu;
        if (NV > 16) nbits |= (hfma2_u(nacc[1], one, k1024) & 0x00ff00ffu) << 8;
    }
}

// One TMEM warp's share of a stage: NC (64, 48 or 32) fp32 columns of S at `base` -> packed fp16 relu(S) IN PLACE over the
// first NC/2 of them (each warp only overwrites columns it has itself loaded, so the two warps of a quarter need no
// ordering between them), sign / near-zero bits of chunk c (32 columns) in sw[c] / nw[c].  The loads are issued separately
// (load_cols) so that the next stage's can be in flight while this stage's masks are being stored.
template <int NC>
__device__ __forceinline__ void load_cols(uint32_t base, uint32_t (&r0)[32], uint32_t (&r1)[32]) {
    static_assert(NC == 64 || NC == 48 || NC == 32, "column split");
    tmem_ld32(base, r0);
    if (NC == 64) tmem_ld32(base + 32, r1);
    if (NC == 48) tmem_ld16(base + 32, reinterpret_cast<uint32_t(&)[16]>(r1));
}
template <int NC, int MODE>
__device__ __forceinline__ void convert_cols(uint32_t base, const uint32_t (&r0)[32], const uint32_t (&r1)[32], uint32_t tau2,
                                             uint32_t (&sw)[2], uint32_t (&nw)[2]) {
    sw[1] = 0;
    nw[1] = 0;
    uint32_t pk[16];
    relu_pack_bits<32, MODE>(r0, pk, tau2, sw[0], nw[0]);
    tmem_st16(base, pk);
    if (NC == 64) {
        relu_pack_bits<32, MODE>(r1, pk, tau2, sw[1], nw[1]);
        tmem_st16(base + 16, pk);
    } else if (NC == 48) {
        relu_pack_bits<16, MODE>(r1, pk, tau2, sw[1], nw[1]);
        tmem_st8(base + 16, reinterpret_cast<uint32_t(&)[8]>(pk));
    }
}

// bits of a mask word (32 columns, bit m = column 2m, bit 16+m = column 2m+1) whose column is below `nvalid`
__device__ __forceinline__ uint32_t valid_bits(int nvalid) {
    if (nvalid >= 32) return 0xffffffffu;
    if (nvalid <= 0) return 0u;
    const uint32_t even = (1u << ((nvalid + 1) >> 1)) - 1u, odd = (1u << (nvalid >> 1)) - 1u;
    return even | (odd << 16);
}

// Segment epilogue of the forward kernel, one TMEM warp's share (rows 32 q .. of the CTA's tile, column half hf).
//   out[n,c,j] = (hi[c] + lo[c]) / (colsum + 1e-5)                     (network/resnet38_contrast.py:72-73)
// A unit swept by ONE CTA pair (every row unit, and every unit when the pair count divides the unit count: nseg == 1) goes
// straight from TMEM to `out` — no partials, no finalize pass (round 1/2 ran a finalize kernel over all units: 10 us + a
// launch on the critical path of both views).  A unit shared by several pairs leaves column-major partials
// ([n][slot][64][HWpj]: for a fixed column the 32 lanes of a warp write one 128-byte line) for pcm_fwd_finalize_kernel.
// The quotient is one reciprocal and a multiply per class in both places.
// NOT inlined, few arguments, few registers: this kernel sits at the 168-register limit and anything that perturbs the
// allocation of the stage loop costs 15-60 % (measured round 3: inlined epilogue 1370 -> 1600 clk per stage; an in-kernel
// "last segment to arrive adds the slots" reduction made the loop spill: 50 -> 110 us).
__device__ __noinline__ void fwd_epilogue(uint32_t o_addr, float* __restrict__ orow, float* __restrict__ dsum_row,
                                          float* __restrict__ part_row, int C, int HW, int HWpj, bool row_live, int hf,
                                          bool whole) {
    if (whole) {
        const int dch = C >> 4;                             // 16-column chunk holding the column-sum column C
        uint32_t r[16], l[16];
        float d = 0.f;
        tmem_ld16(o_addr + dch * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 16; ++v) d = (dch * 16 + v == C) ? __uint_as_float(r[v]) : d;
        d += 1e-5f;
        if (hf == 0) *dsum_row = d;
        if (hf * 16 < C) {                                  // this warp's classes: [16 hf, 16 hf + 16)
            tmem_ld16(o_addr + hf * 16, r);
            tmem_ld16(o_addr + kCamRows + hf * 16, l);
            tmem_ld_wait();
            const float rd = __frcp_rn(d);
            if (row_live) {
#pragma unroll
                for (int v = 0; v < 16; ++v)
                    if (hf * 16 + v < C)
                        orow[static_cast<size_t>(hf * 16 + v) * HW] = (__uint_as_float(r[v]) + __uint_as_float(l[v])) * rd;
            }
        }
        return;
    }
    float* dst = part_row + static_cast<size_t>(hf * 32) * HWpj;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if (hf * 32 + c * 16 >= kONum) break;              // the accumulator has kONum (48 or 64) columns
        uint32_t r[16];
        tmem_ld16(o_addr + hf * 32 + c * 16, r);
        tmem_ld_wait();
#pragma unroll
        for (int v = 0; v < 16; ++v) dst[static_cast<size_t>(c * 16 + v) * HWpj] = __uint_as_float(r[v]);
    }
}

// ---- stream-K schedule shared by the forward and backward main kernels ---------------------------------------------
// Unit u = n*PJ + jp (one image's pair of 128-row j tiles) owns NS stages; the T = units*NS stages are cut into G equal
// contiguous ranges, one per CTA pair, so every pair does the same number of stages whatever N, PJ and NS are (no wave
// quantisation) and pays ONE prologue/drain.  A range crosses unit boundaries: each (pair, unit) piece is a "segment"
// with its own A-side tile and its own partial-result slot  slot = pair - (pair owning the unit's first stage).
// The kernel body: CTA pair `g` of the `G` pairs that share this problem (pcm_fwd_kernel: the whole grid; pcm_fwd_multi_kernel:
// the pairs allotted to one of several problems of different sizes).
template <int WB, int MODE>
__device__ __forceinline__ void
pcm_fwd_body(const uint8_t* __restrict__ fa, const uint8_t* __restrict__ fb, const uint8_t* __restrict__ camb,
             float* __restrict__ part, uint2* __restrict__ mask, uint2* __restrict__ nzmask,
             const float* __restrict__ lo2max, float* __restrict__ out, float* __restrict__ dsum, int C, int K, int HW, int PJ, int NS, int n_units, int nslot,
             long long* __restrict__ trace, const int g, const int G) {
    constexpr int FS = kFStages;
    // optional timeline of CTA 0: trace[role*256 + t*4 + k] = clock64()   (wsegk_debug_pcm_trace)
    const bool tr = trace != nullptr && blockIdx.x == 0;
#define TRACE(role, t, k) do { if (tr && (t) < 64) trace[(role) * 256 + (t) * 4 + (k)] = clock64(); } while (0)
    const bool trp = trace != nullptr && blockIdx.x == 1;          // the peer CTA of pair 0 (its own clock64: same SM pair, clocks differ by a constant)
#define TRACEP(role, t, k) do { if (trp && (t) < 64) trace[(role) * 256 + (t) * 4 + (k)] = clock64(); } while (0)
        // slots 1024 + 2*cta, +1: globaltimer at the start / end of every CTA (the buffer holds 1024 + 2*gridDim.x + 2 words)
    // (CTA 0's pair goes to slots 1024 + 2*gridDim.x ..: the slots of role 0 are all taken by the issuers' stamps)
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt));
        trace[1024 + 2 * blockIdx.x] = static_cast<long long>(gt);
        if (tr) trace[1024 + 2 * gridDim.x] = clock64();
    }
    constexpr int NCOL = 2 * WB;
    constexpr int kSBuf = FwdTmem<WB>::SB;
    constexpr uint32_t kSStride = FwdTmem<WB>::STRIDE, kColO = FwdTmem<WB>::COL_O;
    constexpr uint32_t kCamBytes = NCOL * 2 * kCamRows;        // (NCOL/8) groups x kCamRows rows x 16 B per CTA
    using Bars = FwdBars;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t a_bytes = kTile * K * 2, b_bytes = WB * K * 2;
    uint8_t* s_fj = smem;                                      // 2 x Fj tile: segment k uses slot k & 1, so the next segment's
                                                               // A-side tile is loaded while this one's products still run
    uint8_t* s_f = smem + kABuf * a_bytes;                         // FS x Fi tile
    uint8_t* s_c = s_f + FS * b_bytes;                         // kCamStages x cam block
    Bars* bars = reinterpret_cast<Bars*>(s_c + kCamStages * kCamBytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const PcmSched sched = pcm_sched(n_units, NS, G);
    const StageCursor c0(sched, g, PJ);                        // every role walks its own copy
    const int cnt = c0.cnt;                                    // stages this CTA pair runs
    const int HWpj = PJ * 2 * kTile;

    if (threadIdx.x == 0) {
        const uint32_t relay = rank == 0 ? 1u : 0u;
        for (int s = 0; s < 2; ++s) {
            mbar_init(smem_u32(&bars->fj[s]), 1 + relay);
            mbar_init(smem_u32(&bars->a_free[s]), 1);
        }
        for (int s = 0; s < FS; ++s) {
            mbar_init(smem_u32(&bars->full_f[s]), 1 + relay);
            mbar_init(smem_u32(&bars->empty_f[s]), 1);
        }
        for (int s = 0; s < kCamStages; ++s) {
            mbar_init(smem_u32(&bars->full_c[s]), 1 + relay);
            mbar_init(smem_u32(&bars->empty_c[s]), 1);
        }
        for (int b = 0; b < kSBuf; ++b) {
            mbar_init(smem_u32(&bars->s_full[b]), 1);
            mbar_init(smem_u32(&bars->p_full[b]), kFwdPingPong ? 8 : 16);   // TMEM warps converting a stage x 2 CTAs
        }
        mbar_init(smem_u32(&bars->o_full), 1);
        mbar_init(smem_u32(&bars->o_free), 16);                // 8 TMEM warps x 2 CTAs have drained the accumulator
        mbar_fence_init();
    }
    if (warp == kFwdWarpS) {
        tmem_alloc_2cta(smem_u32(&bars->tmem_base), kTmemCols);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == kFwdWarpProdF) {
        // ---------------------------------------------------------------- feature-tile producer
        if (lane == 0) {
            StageCursor si = c0;
            for (int i = 0; i < cnt; ++i, si.next()) {
                if (si.first()) {
                    // a new segment: its A-side tile goes to slot k & 1, free once every S product of segment k - 2 is done
                    const int as = si.k % kABuf;
                    if (si.k >= kABuf) mbar_wait(smem_u32(&bars->a_free[as]), ((si.k - kABuf) / kABuf) & 1);
                    const size_t a_tile = static_cast<size_t>(si.n) * 2 * PJ + 2 * si.jp + rank;
                    mbar_arrive_expect_tx(smem_u32(&bars->fj[as]), a_bytes);
                    tma_load_1d(smem_u32(s_fj + as * a_bytes), fa + a_tile * a_bytes, a_bytes, smem_u32(&bars->fj[as]));
                }
                const size_t bt = static_cast<size_t>(si.n) * 2 * NS + 2 * si.st + rank;
                const int s = i % FS, u = i / FS;
                mbar_wait(smem_u32(&bars->empty_f[s]), (u & 1) ^ 1);
                TRACE(0, i, 0);
                const uint32_t bar = smem_u32(&bars->full_f[s]);
                mbar_arrive_expect_tx(bar, b_bytes);
                tma_load_1d(smem_u32(s_f + s * b_bytes), fb + bt * b_bytes, b_bytes, bar);
            }
        }
    } else if (warp == kFwdWarpProdC) {
        // ---------------------------------------------------------------- cam-block producer
        if (lane == 0) {
            StageCursor si = c0;
            for (int i = 0; i < cnt; ++i, si.next()) {
                const int s = i % kCamStages, u = i / kCamStages;
                mbar_wait(smem_u32(&bars->empty_c[s]), (u & 1) ^ 1);
                const uint32_t bar = smem_u32(&bars->full_c[s]);
                mbar_arrive_expect_tx(bar, kCamBytes);
                tma_load_1d(smem_u32(s_c + s * kCamBytes),
                            camb + ((static_cast<size_t>(si.n) * NS + si.st) * 2 + rank) * kCamBytes, kCamBytes, bar);
            }
        }
    } else if (warp == kFwdWarpS) {
        if (lane == 0 && rank == 1) {
            // ------------------------------------------------------------ relay: tell the leader my tiles have landed
            StageCursor si = c0;
            for (int i = 0; i < cnt; ++i, si.next()) {
                if (si.first()) {
                    mbar_wait(smem_u32(&bars->fj[si.k % kABuf]), (si.k / kABuf) & 1);
                    mbar_arrive_cluster(mapa_shared(smem_u32(&bars->fj[si.k % kABuf]), 0));
                }
                {
                    const int s = i % FS, u = i / FS;
                    mbar_wait(smem_u32(&bars->full_f[s]), u & 1);
                    mbar_arrive_cluster(mapa_shared(smem_u32(&bars->full_f[s]), 0));
                }
                {
                    const int s = i % kCamStages, u = i / kCamStages;
                    mbar_wait(smem_u32(&bars->full_c[s]), u & 1);
                    mbar_arrive_cluster(mapa_shared(smem_u32(&bars->full_c[s]), 0));
                }
            }
        } else if (rank == 0) {
            // ------------------------------------------------------------ S issuer (leader CTA): S(i) = Fj Fi^T into buffer i % 3.
            // ONE elected lane runs the whole loop inside a single divergent branch: re-electing per stage (elect.sync + the
            // reconvergence around it) cost ~700 clk per stage, and 32 lanes polling an mbarrier ~150 clk per wait even when
            // the phase is complete (tools/ubench/mma_*.cu, trace r2).  The P V products have their own issuing warp: one
            // thread issuing both kinds, with its three waits per stage, left the tensor pipe idle 40 % of the time.
            // Issue itself is cheap (~8 clk per MMA with precomputed descriptors; the queue holds ~24 MMAs).
            if (elect_one()) {
                const uint32_t id_s = umma_idesc_f16(0, 2 * kTile, NCOL, 0, 0);
                constexpr uint32_t lbo_a = kTile * 16, lbo_b = WB * 16;
                const int ksteps = K >> 4;
                const uint64_t da0 = umma_smem_desc(smem_u32(s_fj), lbo_a, 128);
                const uint64_t db0 = umma_smem_desc(smem_u32(s_f), lbo_b, 128);
                StageCursor cs = c0;
                int b = 0, s = 0, fph = 0;
                uint64_t da = da0;
                for (int i = 0; i < cnt; ++i) {
                    const bool seg_first = cs.first(), seg_last = cs.last();
                    const int seg = cs.k;
                    cs.next();
                    if (seg_first) {
                        mbar_wait(smem_u32(&bars->fj[seg % kABuf]), (seg / kABuf) & 1);
                        da = da0 + static_cast<uint64_t>(((seg % kABuf) * a_bytes) >> 4);
                    }
                    TRACE(1, i, 0);
                    // the stage's tile has landed (in both CTAs), and buffer b is free: P V(i - 3) has read it — that product's
                    // commit on empty_c (P V(i + 1), the next completion of the same barrier, needs S(i + 1), which this thread
                    // has not issued yet).  Both polls are in flight together.
                    if (i >= kSBuf)
                        mbar_wait2(smem_u32(&bars->full_f[s]), fph, smem_u32(&bars->empty_c[(i - kSBuf) % kCamStages]),
                                   ((i - kSBuf) / kCamStages) & 1);
                    else
                        mbar_wait(smem_u32(&bars->full_f[s]), fph);
                    TRACE(1, i, 1);
                    tc_fence_after();
                    const uint64_t db = db0 + static_cast<uint64_t>((s * b_bytes) >> 4);
                    const uint32_t d = tmem + b * kSStride;
#ifndef WSEGK_ABL_NOS
#pragma unroll
                    for (int ks = 0; ks < kMaxPcmK / 16; ++ks)
                        if (ks < ksteps)
                            umma2_ss(d, da + static_cast<uint64_t>(ks * (2 * lbo_a / 16)),
                                     db + static_cast<uint64_t>(ks * (2 * lbo_b / 16)), id_s, ks > 0);
#endif
                    TRACE(0, i, 1);
                    umma2_commit(smem_u32(&bars->s_full[b]));
                    umma2_commit(smem_u32(&bars->empty_f[s]));
                    if (seg_last) umma2_commit(smem_u32(&bars->a_free[seg % kABuf]));  // the segment's A-side tile may be replaced
                    TRACE(0, i, 2);
                    b = b + 1 == kSBuf ? 0 : b + 1;
                    if (++s == FS) { s = 0; fph ^= 1; }
                }
            }
            __syncwarp();
        }
    } else if (warp == kFwdWarpPV) {
        if (rank == 0) {
            // ------------------------------------------------------------ P V issuer (leader CTA), one lane, one branch
            if (elect_one()) {
                const uint32_t id_o = umma_idesc_f16(0, 2 * kTile, kONum, 0, 0);
                constexpr uint32_t lbo_c = kCamRows * 16;
                const uint64_t dc0 = umma_smem_desc(smem_u32(s_c), lbo_c, 128);
                StageCursor si = c0;
                int b = 0, bph = 0, sc = 0, cph = 0;
                for (int i = 0; i < cnt; ++i, si.next()) {
                    TRACE(1, i, 2);
                    if (tr) {                                   // diagnosis: which of the two is late
                        mbar_wait(smem_u32(&bars->p_full[b]), bph);
                        TRACE(3, i, 3);
                    }
                    if (kFwdPingPong && si.first() && si.k > 0) mbar_wait(smem_u32(&bars->o_free), (si.k - 1) & 1);
                    mbar_wait2(smem_u32(&bars->p_full[b]), bph, smem_u32(&bars->full_c[sc]), cph);
                    TRACE(1, i, 3);
                    tc_fence_after();
                    const uint64_t dc = dc0 + static_cast<uint64_t>((sc * kCamBytes) >> 4);
                    const uint32_t acc0 = si.first() ? 0u : 1u;
                    // the accumulator restarts with every segment; the TMEM warps drained the previous segment's before
                    // they produced this stage's P.  P of S columns [0,64) sits at packed columns [0,32) of the buffer, P of
                    // the columns from 64 on at [64, ...) (convert_cols)
#ifndef WSEGK_ABL_NOPV
#pragma unroll
                    for (int ks = 0; ks < NCOL / 16; ++ks) {
                        constexpr int kHalf = kSplitCol / 16;
                        const uint32_t pcol = ks < kHalf ? ks * 8 : kSplitCol + (ks - kHalf) * 8;
                        umma2_ts(tmem + kColO, tmem + b * kSStride + pcol, dc + static_cast<uint64_t>(ks * (2 * lbo_c / 16)),
                                 id_o, ks > 0 ? 1u : acc0);
                    }
#endif
                    TRACE(0, i, 3);
                    umma2_commit(smem_u32(&bars->empty_c[sc]));
                    if (si.last()) umma2_commit(smem_u32(&bars->o_full));
                    if (++b == kSBuf) { b = 0; bph ^= 1; }
                    if (++sc == kCamStages) { sc = 0; cph ^= 1; }
                }
            }
            __syncwarp();
        }
    } else if (kFwdPingPong && warp < kFwdTmemWarp0 + 8) {
        // ---------------------------------------------------------------- ReLU / sign / convert warps, alternate stages
        const int q = warp & 3;                         // TMEM lane quarter this warp may touch
        const int par = (warp - kFwdTmemWarp0) >> 2;    // stages j = par (mod 2) are mine
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t p_full0 = mapa_shared(smem_u32(&bars->p_full[0]), 0);
        const uint32_t o_free0 = mapa_shared(smem_u32(&bars->o_free), 0);
        uint32_t tau2 = 0;
        if (MODE == 2) {                                // see the column-split branch below
            const float tau = sqrtf(__ldg(lo2max)) + 1e-6f;
            const uint32_t th = __half_as_ushort(__float2half_ru(tau));
            tau2 = th | (th << 16);
        }
        constexpr int NC1 = NCOL - kSplitCol;           // columns behind the first 64: 64, 48 or 32
        StageCursor si = c0;
        uint32_t r0[32], r1[32];
        auto fetch0 = [&](int j) {                      // wait for S(j) and start loading its first 64 columns
            const int fb = j % kSBuf;
            if (q == 0 && lane == 0) TRACE(2, j, 0);
            mbar_wait(smem_u32(&bars->s_full[fb]), (j / kSBuf) & 1);
            tc_fence_after();
            const uint32_t base = tmem + lane_addr + fb * kSStride;
            tmem_ld32(base, r0);
            tmem_ld32(base + 32, r1);
        };
        if (par < cnt) fetch0(par);
        for (int j = 0; j < cnt; ++j, si.next()) {
            const int row = si.jp * 2 * kTile + static_cast<int>(rank) * kTile + q * 32 + lane;
            if ((j & 1) == par) {
                const int b = j % kSBuf;
                const uint32_t base = tmem + lane_addr + b * kSStride;
                uint32_t sw[4] = {0u, 0u, 0u, 0u}, nw[4] = {0u, 0u, 0u, 0u};
                uint32_t pk[16];
                tmem_ld_wait();
                if (q == 0 && lane == 0) { TRACE(2, j, 1); TRACEP(3, j, 0); }
                // P = relu(S) as packed fp16 IN PLACE: columns [0,64) -> packed columns [0,32), columns [64, NCOL) -> [64, ..);
                // every store lands on columns this warp has already loaded, and the loads of the second part are issued as
                // soon as the registers of the first are free
                relu_pack_bits<32, MODE>(r0, pk, tau2, sw[0], nw[0]);
                tmem_st16(base, pk);
                tmem_ld32(base + kSplitCol, r0);
                relu_pack_bits<32, MODE>(r1, pk, tau2, sw[1], nw[1]);
                tmem_st16(base + 16, pk);
                if (NC1 == 64) tmem_ld32(base + kSplitCol + 32, r1);
                if (NC1 == 48) tmem_ld16(base + kSplitCol + 32, reinterpret_cast<uint32_t(&)[16]>(r1));
                tmem_ld_wait();
                relu_pack_bits<32, MODE>(r0, pk, tau2, sw[2], nw[2]);
                tmem_st16(base + kSplitCol, pk);
                if (NC1 == 64) {
                    relu_pack_bits<32, MODE>(r1, pk, tau2, sw[3], nw[3]);
                    tmem_st16(base + kSplitCol + 16, pk);
                } else if (NC1 == 48) {
                    relu_pack_bits<16, MODE>(r1, pk, tau2, sw[3], nw[3]);
                    tmem_st8(base + kSplitCol + 16, reinterpret_cast<uint32_t(&)[8]>(pk));
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(p_full0 + static_cast<uint32_t>(b) * 8u);
                if (q == 0 && lane == 0) { TRACE(2, j, 2); TRACEP(3, j, 1); }
                if (q == 3 && lane == 0) TRACE(3, j, 2);
                // my next stage's S is normally complete by now: start loading it before this stage's masks are stored
                if (j + 2 < cnt) fetch0(j + 2);
                if (MODE >= 1) {
                    const size_t mi = static_cast<size_t>(si.n * NS + si.st) * HWpj + row;    // [n][stage][row] x 4 words
                    reinterpret_cast<uint4*>(mask)[mi] = make_uint4(sw[0], sw[1], sw[2], sw[3]);
                    if (MODE == 2) {
                        // padding rows / columns are all-zero vectors: S = 0 there, nothing to re-evaluate
                        const int vc = HW - si.st * NCOL;                               // valid columns of this stage
                        if (row >= HW) { nw[0] = 0; nw[1] = 0; nw[2] = 0; nw[3] = 0; }
                        else if (vc < NCOL) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) nw[c] &= valid_bits(vc - 32 * c);
                        }
                        reinterpret_cast<uint4*>(nzmask)[mi] = make_uint4(nw[0], nw[1], nw[2], nw[3]);
                    }
                }
            }
            if (si.last()) {
                // segment epilogue (fwd_epilogue above), both warps of the quarter: half of the classes / columns each.  The
                // P V issuer starts the next segment's accumulation only when all 16 warps of the pair have drained this one
#ifdef WSEGK_FWD_NOFUSE
                const int nseg = 0;
#else
                const int nseg = pcm_unit_segments(sched, si.u);
#endif
                const int slot = (nseg != 1 && si.in_tail())
                                     ? g - pcm_tail_owner(sched, static_cast<long long>(si.u - sched.Rf * sched.G) * NS) : 0;
                mbar_wait(smem_u32(&bars->o_full), si.k & 1);
                tc_fence_after();
                fwd_epilogue(tmem + lane_addr + kColO, out + static_cast<size_t>(si.n) * C * HW + row,
                             dsum + static_cast<size_t>(si.n) * HWpj + row,
                             part + (static_cast<size_t>(si.n) * nslot + slot) * kONum * HWpj + row, C, HW, HWpj, row < HW, par,
                             nseg == 1);
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(o_free0);
            }
        }
    } else if (warp < kFwdTmemWarp0 + 8) {
        // ---------------------------------------------------------------- ReLU / sign / convert warps, column split
        const int q = warp & 3;                         // TMEM lane quarter this warp may touch
        const int hf = (warp - kFwdTmemWarp0) >> 2;     // column half
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t p_full0 = mapa_shared(smem_u32(&bars->p_full[0]), 0);
        uint32_t tau2 = 0;
        if (MODE == 2) {
            // |S(fp16 operands) - S| <= |lo_j| + |lo_i| <= 2 lo_max is the strict bound; the residuals of the K channels add
            // like a random walk, sigma ~ lo_max / 6 (measured: max over 8e7 entries 0.95 lo_max at 5.7 sigma), so an entry
            // beyond tau = lo_max with the wrong sign is a > 6 sigma event that also needs |S| to sit just above tau:
            // ~1e-13 per entry
            const float tau = sqrtf(__ldg(lo2max)) + 1e-6f;
            const uint32_t th = __half_as_ushort(__float2half_ru(tau));
            tau2 = th | (th << 16);
        }
        StageCursor si = c0;
        int b = 0, bph = 0;
        uint32_t r0[32], r1[32];                       // this warp's S columns of the stage being converted / being loaded
        auto fetch = [&](int buf, int phase) {         // wait for S of the stage in `buf` and start loading my columns
            mbar_wait(smem_u32(&bars->s_full[buf]), phase);
            tc_fence_after();
            const uint32_t base = tmem + lane_addr + buf * kSStride + hf * kSplitCol;
            if (hf == 0) load_cols<kSplitCol>(base, r0, r1);
            else load_cols<NCOL - kSplitCol>(base, r0, r1);
        };
        if (warp == kFwdTmemWarp0 && lane == 0) TRACE(2, 0, 0);
        if (cnt > 0) fetch(0, 0);
        for (int i = 0; i < cnt; ++i, si.next()) {
            const int row = si.jp * 2 * kTile + static_cast<int>(rank) * kTile + q * 32 + lane;
            tmem_ld_wait();
            if (warp == kFwdTmemWarp0 && lane == 0) TRACE(2, i, 1);
            if (warp == kFwdTmemWarp0 && lane == 0) TRACEP(3, i, 0);
            uint32_t sw[2], nw[2];
            const uint32_t base = tmem + lane_addr + b * kSStride + hf * kSplitCol;
#if defined(WSEGK_ABL_NOALU)
            tmem_st16(base, reinterpret_cast<uint32_t(&)[16]>(r0));
            sw[0] = r0[0]; sw[1] = r1[0]; nw[0] = r0[1]; nw[1] = r1[1];
#else
            if (hf == 0) convert_cols<kSplitCol, MODE>(base, r0, r1, tau2, sw, nw);
            else convert_cols<NCOL - kSplitCol, MODE>(base, r0, r1, tau2, sw, nw);
#endif
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(p_full0 + static_cast<uint32_t>(b) * 8u);
            if (warp == kFwdTmemWarp0 && lane == 0) TRACE(2, i, 2);
            if (warp == kFwdTmemWarp0 + 7 && lane == 0) TRACE(3, i, 2);
            if (warp == kFwdTmemWarp0 && lane == 0) TRACEP(3, i, 1);
            // the next stage's S is normally complete by now (the S products run up to three stages ahead): start loading it
            // before this stage's masks are stored, so the barrier poll (~150 clk) and the TMEM load latency hide behind them
            const int nb = b + 1 == kSBuf ? 0 : b + 1, nph = b + 1 == kSBuf ? bph ^ 1 : bph;
            if (warp == kFwdTmemWarp0 && lane == 0) TRACE(2, i + 1, 0);
            // (not across a segment's end: the epilogue is a call, and the 64 prefetched values would have to survive it)
            const bool seg_end = si.last();
            if (i + 1 < cnt && !seg_end) fetch(nb, nph);
#if !defined(WSEGK_ABL_NOMASK) && !defined(WSEGK_ABL_NOMASK_LANE0)
            if (MODE >= 1) {
#elif defined(WSEGK_ABL_NOMASK_LANE0)
            if (MODE >= 1 && lane != 0) {
#else
            if (MODE >= 1 && sw[0] == 0x12345u) {
#endif
                const size_t mi = static_cast<size_t>(si.n * NS + si.st) * (2 * HWpj) + 2 * row + hf;
                mask[mi] = make_uint2(sw[0], sw[1]);
                if (MODE == 2) {
                    // padding rows / columns are all-zero vectors: S = 0 there, nothing to re-evaluate
                    const int vc = HW - si.st * NCOL - hf * kSplitCol;          // valid columns of my half in this stage
                    if (row >= HW) { nw[0] = 0; nw[1] = 0; }
                    else if (vc < kSplitCol) { nw[0] &= valid_bits(vc); nw[1] &= valid_bits(vc - 32); }
                    nzmask[mi] = make_uint2(nw[0], nw[1]);
                }
            }
            if (seg_end) {
                // segment epilogue (fwd_epilogue above): the accumulator becomes `out` directly, or a partial slot
#ifdef WSEGK_FWD_NOFUSE
                const int nseg = 0;
#else
                const int nseg = pcm_unit_segments(sched, si.u);
#endif
                const int slot = (nseg != 1 && si.in_tail())
                                     ? g - pcm_tail_owner(sched, static_cast<long long>(si.u - sched.Rf * sched.G) * NS) : 0;
                mbar_wait(smem_u32(&bars->o_full), si.k & 1);
                tc_fence_after();
                fwd_epilogue(tmem + lane_addr + kColO, out + static_cast<size_t>(si.n) * C * HW + row,
                             dsum + static_cast<size_t>(si.n) * HWpj + row,
                             part + (static_cast<size_t>(si.n) * nslot + slot) * kONum * HWpj + row, C, HW, HWpj, row < HW, hf,
                             nseg == 1);
                tc_fence_before();
                if (i + 1 < cnt) fetch(nb, nph);
            }
            b = nb;
            bph = nph;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (trace != nullptr && threadIdx.x == 0) {
        unsigned long long gt;
        asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(gt));
        trace[1024 + 2 * blockIdx.x + 1] = static_cast<long long>(gt);
        if (tr) trace[1024 + 2 * gridDim.x + 1] = clock64();
    }
    cluster_sync_all();
    if (warp == kFwdWarpS) tmem_dealloc_2cta(tmem, kTmemCols);
#undef TRACE
#undef TRACEP
}

template <int WB, int MODE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFwdThreads, 1)
pcm_fwd_kernel(const uint8_t* __restrict__ fa, const uint8_t* __restrict__ fb, const uint8_t* __restrict__ camb,
               float* __restrict__ part, uint2* __restrict__ mask, uint2* __restrict__ nzmask,
               const float* __restrict__ lo2max, float* __restrict__ out, float* __restrict__ dsum, int C, int K, int HW, int PJ, int NS, int n_units, int nslot,
               long long* __restrict__ trace) {
    pcm_fwd_body<WB, MODE>(fa, fb, camb, part, mask, nzmask, lo2max, out, dsum, C, K, HW, PJ, NS, n_units, nslot, trace,
                           static_cast<int>(blockIdx.x >> 1), static_cast<int>(gridDim.x >> 1));
}

// Several forward-only problems of DIFFERENT sizes in ONE launch (the 4 scales x {image, flip} of
// contrast_infer.py:58-71 are 4 problems of batch 2): every problem owns a contiguous range of the grid's CTA pairs, in
// proportion to its stages, and runs the schedule of the single-problem kernel inside it.  Per-scale launches leave the GPU to
// one problem at a time — the 0.5x scale (600 pixels: 30 stages in all) cannot fill it, and every launch pays its own
// prologue and tail.
constexpr int kMaxFwdItems = 8;
struct FwdItem {
    const uint8_t *fa, *fb, *camb;
    float *part, *out, *dsum;
    int HW, PJ, NS, n_units, nslot, g0, G;
};
struct FwdItems {
    int n, C, K;
    FwdItem it[kMaxFwdItems];
};
template <int WB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kFwdThreads, 1)
pcm_fwd_multi_kernel(const __grid_constant__ FwdItems items) {
    const int gp = static_cast<int>(blockIdx.x >> 1);
    int idx = 0;
    for (int i = 1; i < items.n; ++i)
        if (gp >= items.it[i].g0) idx = i;
    const FwdItem& it = items.it[idx];
    pcm_fwd_body<WB, 0>(it.fa, it.fb, it.camb, it.part, nullptr, nullptr, nullptr, it.out, it.dsum, items.C, items.K, it.HW,
                        it.PJ, it.NS, it.n_units, it.nslot, nullptr, gp - it.g0, it.G);
}

// ------------------------------------------------------------------------------------------------
// sign fix-up (precise mode)
// ------------------------------------------------------------------------------------------------
// The forward marked the entries (j, i) with |S_ji| < tau in nzmask; here each is re-evaluated as an fp32 dot product of
// the fp32 normalised features (what the reference's fp32 matmul computes, network/resnet38_contrast.py:71) and its bit of
// the sign mask is set to [S > 0].  S is symmetric and so (up to entries sitting exactly on tau, whose fp16 sign is
// reliable by the definition of tau) is its near-zero mask: only the entries ABOVE the diagonal (i > j) are evaluated
// and each result patches both (j, i) and (i, j) — half the row gathers, which are what this kernel costs (one K-float
// row from L2 per entry).  One CTA = 64 rows j (their features in shared memory, bulk copies) x a chunk of stages;
// marked entries go through a shared-memory queue; a warp evaluates eight entries at a time, four lanes per entry with
// all 12 float4 loads of the lane in flight at once (the first version walked one entry per warp with dependent loads:
// 104 us at N = 8, 56x56).
#ifndef WSEGK_FIX_MINBLOCKS
#define WSEGK_FIX_MINBLOCKS 3
#endif
constexpr int kFixThreads = 256;
constexpr int kFixQueue = 1024;
constexpr int kFixTile = 64;            // rows j per CTA: 48 KB of shared memory at K = 192 -> four CTAs per SM keep the L2 gathers in flight

template <int KG>      // K / 16
__global__ void __launch_bounds__(kFixThreads, WSEGK_FIX_MINBLOCKS) pcm_sign_fixup_kernel(const float* __restrict__ fh32,
                                                                    const uint4* __restrict__ nzmask,
                                                                    uint32_t* __restrict__ mask, int HWpj, int NS,
                                                                    int ncol, int chunk) {
    constexpr int K = 16 * KG;
    extern __shared__ __align__(128) uint8_t fix_smem[];
    float* s_fj = reinterpret_cast<float*>(fix_smem);                          // [kFixTile][K]
    uint32_t* s_q = reinterpret_cast<uint32_t*>(fix_smem + static_cast<size_t>(kFixTile) * K * 4);
    __shared__ int s_cnt;
    __shared__ __align__(8) uint64_t s_bar;
    const int n = blockIdx.z, j0 = blockIdx.x * kFixTile;
    // stages of this CTA: its chunk, minus the stages that lie entirely below the diagonal (columns <= j0)
    const int s_begin = max(static_cast<int>(blockIdx.y) * chunk, j0 / ncol);
    const int s_end = min(NS, (static_cast<int>(blockIdx.y) + 1) * chunk);
    if (s_begin >= s_end) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float* fn = fh32 + static_cast<size_t>(n) * HWpj * K;
    if (tid == 0) {
        s_cnt = 0;
        mbar_init(smem_u32(&s_bar), 1);
        mbar_fence_init();
        const uint32_t bytes = static_cast<uint32_t>(kFixTile) * K * 4;
        mbar_arrive_expect_tx(smem_u32(&s_bar), bytes);
        // two halves: the copy engine streams one operation with only a few requests in flight
        tma_load_1d(smem_u32(s_fj), fn + static_cast<size_t>(j0) * K, bytes / 2, smem_u32(&s_bar));
        tma_load_1d(smem_u32(s_fj) + bytes / 2, fn + static_cast<size_t>(j0) * K + bytes / 8, bytes / 2, smem_u32(&s_bar));
    }
    __syncthreads();
    const int my_row = tid & (kFixTile - 1), my_sub = tid / kFixTile;          // kFixThreads / kFixTile stages are scanned per pass
    const int my_j = j0 + my_row;
    uint32_t pend[4] = {0u, 0u, 0u, 0u};
    int pend_stage = 0;
    int s = s_begin;
    bool any_pend = false, tile_ready = false;
    const int sub = lane >> 2, l4 = lane & 3;                                  // entry of the warp's batch, lane inside it
    while (true) {
        if (!any_pend && s < s_end) {
            const int st = s + my_sub;
            if (st < s_end) {
                const uint4 w = __ldg(nzmask + (static_cast<size_t>(n) * NS + st) * HWpj + my_j);
                pend[0] = w.x; pend[1] = w.y; pend[2] = w.z; pend[3] = w.w;
                pend_stage = st;
                // keep only the entries above the diagonal: column index i = st*ncol + col > j
                const int lim = my_j - st * ncol;                                  // columns <= lim are dropped
                if (lim >= 0) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) pend[c] &= ~valid_bits(lim + 1 - 32 * c);
                }
            }
            s += kFixThreads / kFixTile;
        }
        // push my marked entries: [row:7 | column:7 | stage:18]
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            while (pend[c]) {
                const int slot = atomicAdd(&s_cnt, 1);
                if (slot >= kFixQueue) break;
                const int bit = __ffs(pend[c]) - 1;
                pend[c] &= pend[c] - 1;
                const int col = 32 * c + ((bit & 15) << 1) + (bit >> 4);
                s_q[slot] = (static_cast<uint32_t>(my_row) << 25) | (static_cast<uint32_t>(col) << 18) |
                            static_cast<uint32_t>(pend_stage);
            }
        }
        any_pend = __syncthreads_or((pend[0] | pend[1] | pend[2] | pend[3]) != 0u);
        const int cnt = min(s_cnt, kFixQueue);
        const bool done = !any_pend && s >= s_end;
        if (any_pend || done) {
            if (!tile_ready) {
                mbar_wait(smem_u32(&s_bar), 0);
                tile_ready = true;
            }
            for (int e0 = warp * 8; e0 < cnt; e0 += (kFixThreads / 32) * 8) {
                const bool live = e0 + sub < cnt;
                const uint32_t ent = live ? s_q[e0 + sub] : 0u;
                const int row = ent >> 25, col = (ent >> 18) & 127, st = ent & 0x3ffff;
                const int i = st * ncol + col;
                const float4* fi = reinterpret_cast<const float4*>(fn + static_cast<size_t>(i) * K) + l4;
                const uint32_t fj_s = smem_u32(s_fj + row * K + 4 * l4);
                // a dead lane group reads row 0 (entry 0): every load is unconditional, so all KG of them are in flight at once
                float4 x[KG];
#pragma unroll
                for (int t = 0; t < KG; ++t)      // volatile asm: ptxas otherwise keeps only two of the loads in flight
                    asm volatile("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                                 : "=f"(x[t].x), "=f"(x[t].y), "=f"(x[t].z), "=f"(x[t].w)
                                 : "l"(fi + 4 * t));
                __syncwarp();                     // a scheduling fence: the loads above stay above it
                float acc = 0.f;
#pragma unroll
                for (int t = 0; t < KG; ++t) {
                    float4 y;                     // volatile as well: keeps the shared-memory reads (and the FMAs) behind ALL the loads
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];\n"
                                 : "=f"(y.x), "=f"(y.y), "=f"(y.z), "=f"(y.w)
                                 : "r"(fj_s + 64u * t));
                    acc = fmaf(x[t].x, y.x, acc);
                    acc = fmaf(x[t].y, y.y, acc);
                    acc = fmaf(x[t].z, y.z, acc);
                    acc = fmaf(x[t].w, y.w, acc);
                }
                acc += __shfl_xor_sync(0xffffffffu, acc, 1);
                acc += __shfl_xor_sync(0xffffffffu, acc, 2);
                if (live && l4 < 2) {
                    // l4 = 0 patches (j, i), l4 = 1 the mirrored entry (i, j)
                    const int j = j0 + row;
                    const int pr = l4 == 0 ? j : i, pc = l4 == 0 ? i : j;
                    const int pst = pc / ncol, pcol = pc - pst * ncol;
                    const int b = pcol & 31;
                    const uint32_t bitv = 1u << ((b >> 1) + ((b & 1) << 4));
                    uint32_t* wp = mask + ((static_cast<size_t>(n) * NS + pst) * HWpj + pr) * kMaskWords + (pcol >> 5);
                    if (acc > 0.f) atomicOr(wp, bitv);
                    else atomicAnd(wp, ~bitv);
                }
            }
            __syncthreads();
            if (tid == 0) s_cnt = 0;
            __syncthreads();
            if (done) break;
        }
    }
}

// out[n,c,j] = (sum_slot hi[c] + lo[c]) / (sum_slot colsum + 1e-5)
// for the units that were swept by more than one CTA pair (the others were finished by the main kernel's epilogue; `all`:
// the ablation build, where the main kernel leaves partials for every unit).
// The partials are column-major ([n][slot][64][HWpj]: hi rows 0..C-1, column sum at row C, lo rows 32..32+C-1), so thread
// (c, j) reads and writes along j: no transpose.  One CTA = 32 consecutive rows j x 8 class phases.
__device__ __forceinline__ void pcm_fwd_finalize_body(const float* __restrict__ part, float* __restrict__ out,
                                                      float* __restrict__ dsum, int C, int HW, int HWpj, int nslot, int PJ,
                                                      int NS, int G, int all, int N, int bx, int n) {
    const int lane = threadIdx.x & 31, ph = threadIdx.x >> 5;
    const int j = bx * 32 + lane;                                // HWpj is a multiple of 256
    const int nseg = pcm_unit_segments(pcm_sched(N * PJ, NS, G), n * PJ + j / (2 * kTile));
    if (nseg == 1 && !all) return;                               // uniform over the CTA (32 divides 256)
    const size_t slot_stride = static_cast<size_t>(kONum) * HWpj;
    const float* base = part + static_cast<size_t>(n) * nslot * slot_stride + j;
    float d = 0.f;
    for (int sp = 0; sp < nseg; ++sp) d += __ldg(base + sp * slot_stride + static_cast<size_t>(C) * HWpj);
    d += 1e-5f;
    if (ph == 0) dsum[static_cast<size_t>(n) * HWpj + j] = d;
    if (j >= HW) return;
    for (int c = ph; c < C; c += 8) {
        float hi = 0.f, lo = 0.f;
        for (int sp = 0; sp < nseg; ++sp) {
            hi += __ldg(base + sp * slot_stride + static_cast<size_t>(c) * HWpj);
            lo += __ldg(base + sp * slot_stride + static_cast<size_t>(kCamRows + c) * HWpj);
        }
        out[(static_cast<size_t>(n) * C + c) * HW + j] = (hi + lo) * __frcp_rn(d);
    }
}
__global__ void __launch_bounds__(256) pcm_fwd_finalize_kernel(const float* __restrict__ part,
                                                               float* __restrict__ out,
                                                               float* __restrict__ dsum, int C, int HW,
                                                               int HWpj, int nslot, int PJ, int NS, int G, int all) {
    pcm_fwd_finalize_body(part, out, dsum, C, HW, HWpj, nslot, PJ, NS, G, all, static_cast<int>(gridDim.y), blockIdx.x, blockIdx.y);
}


// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// With G = dL/d(out), g_j = G_j / D_j, delta_j = sum_c out[c,j] G[c,j], d_j = delta_j / D_j:
//   dL/dA_ij = cam_i . g_j - d_j,   dL/dS = dL/dA * [S > 0],   dL/dfh = (dS + dS^T) fh
// and (dS + dS^T)_ji = [S_ji > 0] * (U_j . V_i) with the rank-(2C+2) factors
//   U_j = [ g_j ; cam_j ; -d_j ; 1 ],   V_i = [ cam_i ; g_i ; 1 ; -d_i ]              
// A CTA pair owning 256 j pixels needs, per stage: Tt = Uj Vi^T (K = 48, or 144 with the bf16 hi/lo split), the sign
// mask the forward stored, E = mask .* Tt * 2^e as packed FP16 in TMEM (in place), and dFh_j += E Fi with Fi read
// MN-major (each CTA supplies half of the channels).  No atomics; S is never recomputed.

// U tiles: [128 pixels x Ku] bf16 (A layout); V tiles: [wb pixels x Ku] bf16 (B layout).
// SPLIT: U' = [hi | hi | lo], V' = [hi | lo | hi] so that U'.V' = hi.hi + hi.lo + lo.hi.
template <bool SPLIT>
__global__ void __launch_bounds__(128) pcm_bwd_prep_kernel(const float* __restrict__ cam, const float* __restrict__ out,
                                                           const float* __restrict__ gout,
                                                           const float* __restrict__ dsum, uint4* __restrict__ ua,
                                                           uint4* __restrict__ vb, unsigned int* __restrict__ bound,
                                                           int C, int HW, int PJ, int NS, int wb) {
    __shared__ float s_u[kAugK][kTile];
    __shared__ float s_v[kAugK][kTile];
    const int t = blockIdx.x, n = blockIdx.y;
    const int px = threadIdx.x;
    const int p = t * kTile + px;
    const int HWpj = PJ * 2 * kTile, HWpi = NS * 2 * wb;
    for (int i = 0; i < kAugK; ++i) { s_u[i][px] = 0.f; s_v[i][px] = 0.f; }
    float bnd = 0.f;
    if (p < HW) {
        const float invd = 1.f / dsum[static_cast<size_t>(n) * HWpj + p];
        const size_t base = static_cast<size_t>(n) * C * HW + p;
        float delta = 0.f;
        for (int c0 = 0; c0 < C; c0 += 8) {                   // 24 independent loads in flight
            float cm[8], g[8], o[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const bool ok = c0 + j < C;
                const size_t off = base + static_cast<size_t>(c0 + j) * HW;
                cm[j] = ok ? __ldg(cam + off) : 0.f;
                g[j] = ok ? __ldg(gout + off) : 0.f;
                o[j] = ok ? __ldg(out + off) : 0.f;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = c0 + j;
                if (c < C) {
                    delta = fmaf(o[j], g[j], delta);
                    const float gs = g[j] * invd;
                    bnd += fabsf(gs);
                    s_u[c][px] = gs;        s_v[c][px] = cm[j];
                    s_u[C + c][px] = cm[j]; s_v[C + c][px] = gs;
                }
            }
        }
        const float dj = delta * invd;
        bnd += fabsf(dj);
        s_u[2 * C][px] = -dj;     s_v[2 * C][px] = 1.f;
        s_u[2 * C + 1][px] = 1.f; s_v[2 * C + 1][px] = -dj;
    }
    // |T_ji| <= 2 * max_p (sum_c |g_p[c]| + |d_p|) since 0 <= cam <= 1
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) bnd = fmaxf(bnd, __shfl_xor_sync(0xffffffffu, bnd, o));
    if ((threadIdx.x & 31) == 0 && bnd > 0.f)
        atomicMax(bound, __float_as_uint(2.f * bnd));       // non-negative floats order like their bit patterns
    // each thread only reads back its own column of s_u / s_v
    constexpr int GU = (SPLIT ? 3 : 1) * (kAugK / 8);
    const bool in_a = p < HWpj, in_b = p < HWpi;
    const int tb = p / wb, rb = p - tb * wb;
    uint4* du = ua + (static_cast<size_t>(n) * 2 * PJ + t) * GU * kTile + px;
    uint4* dv = vb + (static_cast<size_t>(n) * 2 * NS + tb) * GU * wb + rb;
#pragma unroll
    for (int g = 0; g < kAugK / 8; ++g) {
        uint32_t uh[4], ul[4], vh[4], vl[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float u0 = s_u[8 * g + 2 * e][px], u1 = s_u[8 * g + 2 * e + 1][px];
            const float v0 = s_v[8 * g + 2 * e][px], v1 = s_v[8 * g + 2 * e + 1][px];
            const float u0h = __bfloat162float(__float2bfloat16_rn(u0)), u1h = __bfloat162float(__float2bfloat16_rn(u1));
            const float v0h = __bfloat162float(__float2bfloat16_rn(v0)), v1h = __bfloat162float(__float2bfloat16_rn(v1));
            uh[e] = pack_bf16x2(u0h, u1h);  ul[e] = pack_bf16x2(u0 - u0h, u1 - u1h);
            vh[e] = pack_bf16x2(v0h, v1h);  vl[e] = pack_bf16x2(v0 - v0h, v1 - v1h);
        }
        const uint4 UH = make_uint4(uh[0], uh[1], uh[2], uh[3]), UL = make_uint4(ul[0], ul[1], ul[2], ul[3]);
        const uint4 VH = make_uint4(vh[0], vh[1], vh[2], vh[3]), VL = make_uint4(vl[0], vl[1], vl[2], vl[3]);
        if (in_a) {
            du[static_cast<size_t>(g) * kTile] = UH;
            if (SPLIT) {
                du[static_cast<size_t>(kAugK / 8 + g) * kTile] = UH;
                du[static_cast<size_t>(2 * (kAugK / 8) + g) * kTile] = UL;
            }
        }
        if (in_b) {
            dv[static_cast<size_t>(g) * wb] = VH;
            if (SPLIT) {
                dv[static_cast<size_t>(kAugK / 8 + g) * wb] = VL;
                dv[static_cast<size_t>(2 * (kAugK / 8) + g) * wb] = VH;
            }
        }
    }
}

// power-of-two scale that brings |T| below 2^14 so E = T * scale is representable in FP16 with full precision
__device__ __forceinline__ float pcm_grad_scale(const float* bound) {
    const float b = *bound;
    if (!(b > 0.f) || !isfinite(b)) return 1.f;
    int e;
    frexpf(b, &e);                       // b = m * 2^e, 0.5 <= m < 1  =>  b < 2^e
    return exp2f(static_cast<float>(14 - e));
}

constexpr int kBwdStages = 4;

struct BwdBars {
    uint64_t uj, u_free, full_v[kBwdStages], empty_v[kBwdStages], full_f[kBwdStages], empty_f[kBwdStages], t_full[3],
        e_full[3], g_full;
    uint32_t tmem_base;
};
// backward: T / E buffers in TMEM.  The loop T(i) -> convert -> G(i) issued -> T(i + NT) takes ~2900 clk (trace r3) against
// 1283 clk of tensor-pipe work per stage with two buffers; THREE buffers make it pipe-bound, and three fit beside the 192
// accumulator columns only for 48-row tiles (3 x 96 + 192 = 480 <= 512; 56-row tiles would need 528)
template <int WB>
struct BwdTmem {
#ifdef WSEGK_BWD_TBUF2
    static constexpr int NT = 2;
#else
    static constexpr int NT = (3 * 2 * WB + kMaxPcmK <= 512) ? 3 : 2;
#endif
    static constexpr uint32_t STRIDE = NT == 3 ? 2 * WB : 128;
    static constexpr uint32_t COL_G = NT == 3 ? 320 : 256;
    static_assert(NT * STRIDE <= COL_G && COL_G + kMaxPcmK <= 512, "TMEM columns");
};

// NV values of T (fp32) -> E = [sign bit set] * T * scale as packed fp16.  The sign word has bit m = column 2m, bit 16+m =
// column 2m+1.  Per pair: two FMUL and one IMAD on the FMA pipe (the 0 / 0xffff half masks come from a multiply), the
// convert, one shift and two LOP3 on the ALU pipe — the predicated-select form the compiler made of the obvious code put
// five ALU-pipe instructions on every pair, and the ALU pipe is what bounds a TMEM warp (tools/ubench/alu_rates.cu).
template <int NV>
__device__ __forceinline__ void mask_scale_pack(const uint32_t* r, uint32_t bits, float scale, uint32_t* pk) {
#pragma unroll
    for (int m = 0; m < NV / 2; ++m) {
        const uint32_t h = pack_f16x2(__uint_as_float(r[2 * m]) * scale, __uint_as_float(r[2 * m + 1]) * scale);
        const uint32_t keep = ((bits >> m) & 0x00010001u) * 0xffffu;
        pk[m] = h & keep;
    }
}

template <int WB, bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kMainThreads, 1)
pcm_bwd_kernel(const uint8_t* __restrict__ ua, const uint8_t* __restrict__ vb, const uint8_t* __restrict__ fb,
               const uint4* __restrict__ mask, const float* __restrict__ bound, float* __restrict__ dfhat, int K,
               int PJ, int NS, int n_units, int nslot, long long* __restrict__ trace) {
    // optional timeline of CTA 0 (wsegk_debug_pcm_trace): trace[role*256 + t*4 + k] = clock64()
    const bool tr = trace != nullptr && blockIdx.x == 0;
#define TRACE(role, t, k) do { if (tr && (t) < 64) trace[(role) * 256 + (t) * 4 + (k)] = clock64(); } while (0)
    constexpr int NCOL = 2 * WB;
    constexpr int NT = BwdTmem<WB>::NT;
    constexpr uint32_t kTStride = BwdTmem<WB>::STRIDE, kColG = BwdTmem<WB>::COL_G;
    constexpr int KU = SPLIT ? 3 * kAugK : kAugK;
    constexpr uint32_t kUBytes = kTile * KU * 2, kVBytes = WB * KU * 2;
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t f_part = (K >> 4) * WB * 16;                // one tile's channel half: K/16 groups x WB rows x 16 B
    const uint32_t f_stage = 2 * f_part;
    uint8_t* s_u = smem;
    uint8_t* s_v = smem + kUBytes;
    uint8_t* s_f = s_v + kBwdStages * kVBytes;
    BwdBars* bars = reinterpret_cast<BwdBars*>(s_f + kBwdStages * f_stage);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    // the same schedule as the forward (common.cuh)
    const int g = static_cast<int>(blockIdx.x >> 1);
    const PcmSched sched = pcm_sched(n_units, NS, static_cast<int>(gridDim.x >> 1));
    const StageCursor c0(sched, g, PJ);
    const int cnt = c0.cnt;
    const int HWpj = PJ * 2 * kTile;

    if (threadIdx.x == 0) {
        const uint32_t relay = rank == 0 ? 1u : 0u;
        mbar_init(smem_u32(&bars->uj), 1 + relay);
        mbar_init(smem_u32(&bars->u_free), 1);
        for (int s = 0; s < kBwdStages; ++s) {
            mbar_init(smem_u32(&bars->full_v[s]), 1 + relay);
            mbar_init(smem_u32(&bars->empty_v[s]), 1);
            mbar_init(smem_u32(&bars->full_f[s]), 2 + relay);        // two producer warps (X part, Y part)
            mbar_init(smem_u32(&bars->empty_f[s]), 1);
        }
        for (int b = 0; b < NT; ++b) {
            mbar_init(smem_u32(&bars->t_full[b]), 1);
            // the TMEM warps that own columns of a stage, in both CTAs
            mbar_init(smem_u32(&bars->e_full[b]), kBwdTW == 4 ? 2 * 4 * ((NCOL + 31) / 32) : 16);
        }
        mbar_init(smem_u32(&bars->g_full), 1);
        mbar_fence_init();
    }
    if (warp == kBwdWarpIssue) {
        tmem_alloc_2cta(smem_u32(&bars->tmem_base), kTmemCols);
        tmem_relinquish_2cta();
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;
    const size_t b_bytes = static_cast<size_t>(WB) * K * 2;

    if (warp == kBwdWarpProdUV) {
        // ---------------------------------------------------------------- producer A: the segment's U tile and the V tiles
        if (lane == 0) {
            StageCursor si = c0;
            for (int i = 0; i < cnt; ++i, si.next()) {
                if (si.first()) {
                    if (si.k > 0) mbar_wait(smem_u32(&bars->u_free), (si.k - 1) & 1);
                    const size_t a_tile = static_cast<size_t>(si.n) * 2 * PJ + 2 * si.jp + rank;
                    mbar_arrive_expect_tx(smem_u32(&bars->uj), kUBytes);
                    tma_load_1d(smem_u32(s_u), ua + a_tile * kUBytes, kUBytes, smem_u32(&bars->uj));
                }
                const int s = i % kBwdStages, u = i / kBwdStages;
                mbar_wait(smem_u32(&bars->empty_v[s]), (u & 1) ^ 1);
                TRACE(0, i, 0);
                const uint32_t bar = smem_u32(&bars->full_v[s]);
                mbar_arrive_expect_tx(bar, kVBytes);
                tma_load_1d(smem_u32(s_v + s * kVBytes),
                            vb + (static_cast<size_t>(si.n) * 2 * NS + 2 * si.st + rank) * kVBytes, kVBytes, bar);
            }
        }
    } else if (warp >= kBwdWarpProdX) {
        // ---------------------------------------------------------------- producers B, C: my half of the channels (groups
        // [rank*K/16, (rank+1)*K/16)) of the stage's X tile (warp 6) and Y tile (warp 7)
        if (lane == 0) {
            const int part = warp - kBwdWarpProdX;
            StageCursor si = c0;
            for (int i = 0; i < cnt; ++i, si.next()) {
                const int s = i % kBwdStages, u = i / kBwdStages;
                mbar_wait(smem_u32(&bars->empty_f[s]), (u & 1) ^ 1);
                if (part == 0) TRACE(0, i, 1);
                const uint32_t bar = smem_u32(&bars->full_f[s]);
                mbar_arrive_expect_tx(bar, f_part);
                const uint8_t* x = fb + (static_cast<size_t>(si.n) * 2 * NS + 2 * si.st + part) * b_bytes +
                                   static_cast<size_t>(rank) * f_part;
                tma_load_1d(smem_u32(s_f + s * f_stage + part * f_part), x, f_part, bar);
            }
        }
    } else if (warp == kBwdWarpIssue) {
        if (lane == 0 && rank == 1) {
            StageCursor si = c0;
            for (int i = 0; i < cnt; ++i, si.next()) {
                if (si.first()) {
                    mbar_wait(smem_u32(&bars->uj), si.k & 1);
                    mbar_arrive_cluster(mapa_shared(smem_u32(&bars->uj), 0));
                }
                const int s = i % kBwdStages, u = i / kBwdStages;
                mbar_wait(smem_u32(&bars->full_v[s]), u & 1);
                mbar_arrive_cluster(mapa_shared(smem_u32(&bars->full_v[s]), 0));
                mbar_wait(smem_u32(&bars->full_f[s]), u & 1);
                mbar_arrive_cluster(mapa_shared(smem_u32(&bars->full_f[s]), 0));
            }
        } else if (rank == 0) {
            // ONE elected lane runs the whole issue loop inside a single branch (see the forward kernel: re-electing per stage
            // and polling with 32 lanes cost the issuer more than the tensor pipe needs for the stage's products)
            if (elect_one()) {
                const uint32_t id_t = umma_idesc_f16(1, 2 * kTile, NCOL, 0, 0);       // bf16 factors
                const uint32_t id_g = umma_idesc_f16(0, 2 * kTile, K, 0, 1);          // fp16, B = Fi read MN-major (N = channel)
                constexpr uint32_t lbo_a = kTile * 16, lbo_b = WB * 16;
                const uint64_t du = umma_smem_desc(smem_u32(s_u), lbo_a, 128);
                const uint64_t dv0 = umma_smem_desc(smem_u32(s_v), lbo_b, 128);
                StageCursor cs = c0;                 // cursor of the T products (ahead of the G products)
                auto issue_t = [&](int i) {
                    const bool seg_first = cs.first(), seg_last = cs.last();
                    const int seg = cs.k;
                    cs.next();
                    const int s = i % kBwdStages, u = i / kBwdStages, b = i % NT;
                    if (seg_first) mbar_wait(smem_u32(&bars->uj), seg & 1);
                    TRACE(1, i, 0);
                    mbar_wait(smem_u32(&bars->full_v[s]), u & 1);
                    TRACE(1, i, 1);
                    tc_fence_after();
                    const uint64_t dv = dv0 + static_cast<uint64_t>((s * kVBytes) >> 4);
#ifndef WSEGK_ABL_NOT
#pragma unroll
                    for (int ks = 0; ks < KU / 16; ++ks)
                        umma2_ss(tmem + b * kTStride, du + static_cast<uint64_t>(ks * (2 * lbo_a / 16)),
                                 dv + static_cast<uint64_t>(ks * (2 * lbo_b / 16)), id_t, ks > 0);
#endif
                    umma2_commit(smem_u32(&bars->t_full[b]));
                    umma2_commit(smem_u32(&bars->empty_v[s]));
                    if (seg_last) umma2_commit(smem_u32(&bars->u_free));
                };
                // NT stages of lookahead inside a segment (T(i + NT) overwrites the buffer G(i) reads: issued behind it by this
                // same thread, the tensor pipe keeps the order), one less across a segment boundary (see the forward kernel)
                int next_t = 0;
                auto issue_upto = [&](int limit) {
                    while (next_t < cnt && next_t <= limit + NT) {
                        if (next_t == limit + NT && next_t > 0 && cs.first()) break;
                        issue_t(next_t);
                        ++next_t;
                    }
                };
                issue_upto(-1);
                StageCursor si = c0;
                for (int i = 0; i < cnt; ++i, si.next()) {
                    const int s = i % kBwdStages, u = i / kBwdStages, b = i % NT;
                    TRACE(1, i, 2);
                    if (tr) {
                        mbar_wait(smem_u32(&bars->e_full[b]), (i / NT) & 1);
                        TRACE(0, i, 2);
                    }
                    mbar_wait2(smem_u32(&bars->e_full[b]), (i / NT) & 1, smem_u32(&bars->full_f[s]), u & 1);
                    TRACE(1, i, 3);
                    tc_fence_after();
                    const uint32_t f0 = smem_u32(s_f + s * f_stage);
                    const uint32_t acc0 = si.first() ? 0u : 1u;
                    // dFh[j, ch] += E[j, q] * F[q, ch] over the stage's pixels q = [X(0:WB) | Y(0:WB)]; B is N(ch) x K(q),
                    // MN-major: LBO = step between the two 8-pixel groups of the k-step, SBO = channel-group step
#ifndef WSEGK_ABL_NOG
#pragma unroll
                    for (int ks = 0; ks < NCOL / 16; ++ks) {
                        const int q0 = 16 * ks, q1 = q0 + 8;
                        const uint32_t a0 = f0 + (q0 / WB) * f_part + (q0 % WB) * 16;
                        const uint32_t a1 = f0 + (q1 / WB) * f_part + (q1 % WB) * 16;
                        // E of the T columns [0,64) sits at packed columns [0,32) of the buffer, E of the columns from 64 on at
                        // [64, ...) (the two TMEM warps of a quarter each overwrite only what they loaded)
                        // (TW = 4: a warp owns 32 T columns and leaves their E at the first 16 columns of its quarter)
                        constexpr int kHalf = kSplitCol / 16;
                        const uint32_t ecol = kBwdTW == 4 ? static_cast<uint32_t>((ks >> 1) * 32 + (ks & 1) * 8)
                                                          : (ks < kHalf ? ks * 8 : kSplitCol + (ks - kHalf) * 8);
                        umma2_ts(tmem + kColG, tmem + b * kTStride + ecol, umma_smem_desc(a0, a1 - a0, lbo_b), id_g,
                                 ks > 0 ? 1u : acc0);
                    }
#endif
                    umma2_commit(smem_u32(&bars->empty_f[s]));
                    if (si.last()) umma2_commit(smem_u32(&bars->g_full));
                    issue_upto(i);
                }
            }
            __syncwarp();
        }
    } else if (kBwdTW == 4 && warp < kBwdTmemWarps) {
        // ---------------------------------------------------------------- TMEM warps, four per lane quarter: warp (q, cq)
        // owns the T columns [32 cq, 32 cq + 32) of its 32 rows (the last chunk of a 112-column stage has 16, of a 96-column
        // stage none: that warp only helps with the epilogue) and one sign word per row and stage
        const int q = warp & 3, cq = warp >> 2;
        const int cw = NCOL - 32 * cq >= 32 ? 32 : (NCOL - 32 * cq > 0 ? NCOL - 32 * cq : 0);
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int row_in = static_cast<int>(rank) * kTile + q * 32 + lane;
        const uint32_t e_full0 = mapa_shared(smem_u32(&bars->e_full[0]), 0);
        const float scale = pcm_grad_scale(bound);
        StageCursor si = c0, sn = c0;      // this stage / the next one (mask prefetch)
        auto mask_of = [&](const StageCursor& c) {       // my sign word of the row
            return __ldg(reinterpret_cast<const uint32_t*>(mask) +
                         ((static_cast<size_t>(c.n) * NS + c.st) * HWpj + c.jp * 2 * kTile + row_in) * kMaskWords + cq);
        };
        uint32_t r0[32];
        auto fetch = [&](int i) {
            if (warp == 0 && lane == 0) TRACE(2, i, 0);
            mbar_wait(smem_u32(&bars->t_full[i % NT]), (i / NT) & 1);
            tc_fence_after();
            const uint32_t base = tmem + lane_addr + (i % NT) * kTStride + 32 * cq;
            if (cw == 32) tmem_ld32(base, r0);
            else tmem_ld16(base, reinterpret_cast<uint32_t(&)[16]>(r0));
        };
        uint32_t mw_next = (cnt > 0 && cw > 0) ? mask_of(sn) : 0u;
        if (cnt > 0 && cw > 0) fetch(0);
        for (int i = 0; i < cnt; ++i, si.next()) {
            if (cw > 0) {
                const int b = i % NT;
                const uint32_t mw = mw_next;
                tmem_ld_wait();
                if (warp == 0 && lane == 0) TRACE(2, i, 1);
                const uint32_t base = tmem + lane_addr + b * kTStride + 32 * cq;
                uint32_t pk[16];
                if (cw == 32) {
                    mask_scale_pack<32>(r0, mw, scale, pk);
                    tmem_st16(base, pk);
                } else {
                    mask_scale_pack<16>(r0, mw, scale, pk);
                    tmem_st8(base, reinterpret_cast<uint32_t(&)[8]>(pk));
                }
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster_relaxed(e_full0 + static_cast<uint32_t>(b) * 8u);
                if (warp == 0 && lane == 0) TRACE(2, i, 2);
                sn.next();
                if (i + 1 < cnt) mw_next = mask_of(sn);
                if (i + 1 < cnt) fetch(i + 1);
            }
            if (si.last()) {
                // segment epilogue.  dfhat [n][slot][K][HWpj]: for a fixed channel the 32 lanes of a warp write 32
                // consecutive pixels; the four warps of a quarter take a quarter of the channel groups each
                const int slot = si.in_tail() ? g - pcm_tail_owner(sched, static_cast<long long>(si.u - sched.Rf * sched.G) * NS) : 0;
                mbar_wait(smem_u32(&bars->g_full), si.k & 1);
                tc_fence_after();
                float* dst = dfhat + ((static_cast<size_t>(si.n) * nslot + slot) * K) * HWpj + si.jp * 2 * kTile + row_in;
                const int cgroups = K >> 4, per = (cgroups + 3) >> 2;
                const int c_begin = cq * per, c_end = min(cgroups, c_begin + per);
                for (int c = c_begin; c < c_end; ++c) {
                    uint32_t ra[16];
                    tmem_ld16(tmem + lane_addr + kColG + c * 16, ra);
                    tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < 16; ++v) dst[static_cast<size_t>(c * 16 + v) * HWpj] = __uint_as_float(ra[v]);
                }
                tc_fence_before();
            }
        }
    } else if (warp < kBwdTmemWarps) {
        // ---------------------------------------------------------------- TMEM warps: two per lane quarter, the first takes the
        // T columns [0,64) of its 32 rows, the second the rest (as in the forward kernel); each loads the NEXT stage's columns
        // as soon as it has handed this stage's E over, so the barrier poll and the TMEM load latency overlap
        const int q = warp & 3, hf = warp >> 2;
        constexpr int NC1 = NCOL - kSplitCol;            // 64, 48 or 32 columns for the second warp
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        const int row_in = static_cast<int>(rank) * kTile + q * 32 + lane;
        const uint32_t e_full0 = mapa_shared(smem_u32(&bars->e_full[0]), 0);
        const float scale = pcm_grad_scale(bound);
        StageCursor si = c0, sn = c0;      // this stage / the next one (mask prefetch)
        const uint2* mask2 = reinterpret_cast<const uint2*>(mask);
        auto mask_of = [&](const StageCursor& c) {       // my two sign words of the row
            return __ldg(mask2 + ((static_cast<size_t>(c.n) * NS + c.st) * HWpj + c.jp * 2 * kTile + row_in) * 2 + hf);
        };
        uint32_t r0[32], r1[32];
        auto fetch = [&](int i) {
            if (warp == 0 && lane == 0) TRACE(2, i, 0);
            mbar_wait(smem_u32(&bars->t_full[i % NT]), (i / NT) & 1);
            tc_fence_after();
            const uint32_t base = tmem + lane_addr + (i % NT) * kTStride + hf * kSplitCol;
            if (hf == 0) load_cols<kSplitCol>(base, r0, r1);
            else load_cols<NC1>(base, r0, r1);
        };
        uint2 mw_next = cnt > 0 ? mask_of(sn) : make_uint2(0, 0);
        if (cnt > 0) fetch(0);
        for (int i = 0; i < cnt; ++i, si.next()) {
            const int b = i % NT;
            const uint2 mw = mw_next;
            tmem_ld_wait();
            if (warp == 0 && lane == 0) TRACE(2, i, 1);
            const uint32_t base = tmem + lane_addr + b * kTStride + hf * kSplitCol;
            uint32_t pk[16];
            mask_scale_pack<32>(r0, mw.x, scale, pk);
            tmem_st16(base, pk);
            if (hf == 0 || NC1 == 64) {
                mask_scale_pack<32>(r1, mw.y, scale, pk);
                tmem_st16(base + 16, pk);
            } else if (NC1 == 48) {
                mask_scale_pack<16>(r1, mw.y, scale, pk);
                tmem_st8(base + 16, reinterpret_cast<uint32_t(&)[8]>(pk));
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster_relaxed(e_full0 + static_cast<uint32_t>(b) * 8u);
            if (warp == 0 && lane == 0) TRACE(2, i, 2);
            // the next stage's sign words are requested only now: an arrive is held back behind the warp's outstanding
            // global accesses (trace r2 of the forward kernel), and "E is ready" is on the critical path
            sn.next();
            if (i + 1 < cnt) mw_next = mask_of(sn);
            if (i + 1 < cnt) fetch(i + 1);
            if (si.last()) {
                // segment epilogue.  dfhat [n][slot][K][HWpj]: for a fixed channel the 32 lanes of a warp write 32
                // consecutive pixels; the two warps of a quarter take half of the channels each
                const int slot = si.in_tail() ? g - pcm_tail_owner(sched, static_cast<long long>(si.u - sched.Rf * sched.G) * NS) : 0;
                mbar_wait(smem_u32(&bars->g_full), si.k & 1);
                tc_fence_after();
                float* dst = dfhat + ((static_cast<size_t>(si.n) * nslot + slot) * K) * HWpj + si.jp * 2 * kTile + row_in;
                const int cgroups = K >> 4, c_begin = hf * ((cgroups + 1) >> 1), c_end = hf ? cgroups : ((cgroups + 1) >> 1);
                // two 16-column chunks in flight: the next chunk's TMEM load is issued before this chunk's stores (the drain
                // is latency-bound — trace r3: ~5000 clk per segment with one load -> wait -> 16 stores at a time — and the
                // G products of the next segment wait for it)
                uint32_t ra[16], rb[16];
                auto put = [&](int c, const uint32_t (&r)[16]) {
#pragma unroll
                    for (int v = 0; v < 16; ++v) dst[static_cast<size_t>(c * 16 + v) * HWpj] = __uint_as_float(r[v]);
                };
                if (c_begin < c_end) tmem_ld16(tmem + lane_addr + kColG + c_begin * 16, ra);
                for (int c = c_begin; c < c_end; c += 2) {
                    tmem_ld_wait();
                    if (c + 1 < c_end) tmem_ld16(tmem + lane_addr + kColG + (c + 1) * 16, rb);
                    put(c, ra);
                    if (c + 1 < c_end) {
                        tmem_ld_wait();
                        if (c + 2 < c_end) tmem_ld16(tmem + lane_addr + kColG + (c + 2) * 16, ra);
                        put(c + 1, rb);
                    }
                }
                tc_fence_before();
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    if (warp == kBwdWarpIssue) tmem_dealloc_2cta(tmem, kTmemCols);
#undef TRACE
}

// grad_f[n,k,p] = dfh[k] * inv - f[k] * (sum_k dfh[k] f[k]) * inv^2 / |f|,  inv = 1/(|f| + 1e-5)
// (autograd of f / (norm(f) + 1e-5), network/resnet38_contrast.py:70); dfh = sum over the i-splits / 2^e.
// CTA = 32 pixels x 8 channel phases; every global access is a coalesced 128-byte row segment.  Pass 1 computes the
// per-pixel dot product, pass 2 re-reads (L2) and writes; loads are issued in batches of 8 independent requests.
constexpr int kPostPix = 32;
constexpr int kPostPhases = 8;
__global__ void __launch_bounds__(256) pcm_bwd_post_kernel(const float* __restrict__ dfhat, const float* __restrict__ f,
                                                           const float* __restrict__ inv_norm,
                                                           const float* __restrict__ bound, float* __restrict__ grad_f,
                                                           int K, int HW, int HWpj, int nslot, int PJ, int NS, int G) {
    __shared__ float s_dot[kPostPhases][kPostPix];
    const int n = blockIdx.y;
    const int px = threadIdx.x & (kPostPix - 1), ph = threadIdx.x >> 5;
    const int p = blockIdx.x * kPostPix + px;
    const bool live = p < HW;
    const float unscale = 1.f / pcm_grad_scale(bound);
    const size_t split_stride = static_cast<size_t>(K) * HWpj;
    const float* gsrc = dfhat + static_cast<size_t>(n) * nslot * split_stride + p;
    // how many CTA pairs shared this pixel's (image, j pair) sweep (kPostPix divides 256: uniform over the CTA)
    const int nsplit = pcm_unit_segments(pcm_sched(static_cast<int>(gridDim.y) * PJ, NS, G),
                                         n * PJ + (blockIdx.x * kPostPix) / (2 * kTile));
    const float* fsrc = f + static_cast<size_t>(n) * K * HW + p;
    const int per = K / kPostPhases;                      // channels per phase (K % 32 == 0 => per % 4 == 0)
    const int k0 = ph * per;
    float dot = 0.f;
    if (live) {
        for (int kk = 0; kk < per; kk += 4) {
            float g[4] = {0.f, 0.f, 0.f, 0.f}, fv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) fv[j] = __ldg(fsrc + static_cast<size_t>(k0 + kk + j) * HW);
            if (nsplit == 2) {                             // the common case: all 12 loads of the batch independent
                float g2[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    g[j] = __ldg(gsrc + static_cast<size_t>(k0 + kk + j) * HWpj);
                    g2[j] = __ldg(gsrc + split_stride + static_cast<size_t>(k0 + kk + j) * HWpj);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) g[j] += g2[j];
            } else {
                for (int sp = 0; sp < nsplit; ++sp) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) g[j] += __ldg(gsrc + sp * split_stride + static_cast<size_t>(k0 + kk + j) * HWpj);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) dot = fmaf(g[j] * unscale, fv[j], dot);
        }
    }
    s_dot[ph][px] = dot;
    __syncthreads();
    if (!live) return;
    dot = 0.f;
#pragma unroll
    for (int i = 0; i < kPostPhases; ++i) dot += s_dot[i][px];
    const float inv = inv_norm[static_cast<size_t>(n) * HWpj + p];
    const float nrm = 1.f / inv - 1e-5f;
    const float coef = nrm > 0.f ? dot * inv * inv / nrm : 0.f;
    float* dst = grad_f + static_cast<size_t>(n) * K * HW + p;
    for (int kk = 0; kk < per; kk += 4) {
        float g[4] = {0.f, 0.f, 0.f, 0.f}, fv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) fv[j] = __ldg(fsrc + static_cast<size_t>(k0 + kk + j) * HW);
        for (int sp = 0; sp < nsplit; ++sp) {
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] += __ldg(gsrc + sp * split_stride + static_cast<size_t>(k0 + kk + j) * HWpj);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[static_cast<size_t>(k0 + kk + j) * HW] = g[j] * unscale * inv - fv[j] * coef;
    }
}

// The same in ONE pass for HW % 4 == 0: a CTA = 32 pixels (8 float4 lanes) x 32 channel groups of PER = K/32 channels; a
// thread keeps its PER x 4 gradient and feature values in registers between the dot product and the write, and has all of
// its 16-byte loads (PER x (1 + segments)) in flight at once.  For a fixed channel 8 lanes read / write one 128-byte line.
// (The two-pass kernel above re-read dfhat from L2 and held 12 four-byte loads per thread: 39 us at N = 8, 56x56.)
template <int PER>
__global__ void __launch_bounds__(256) pcm_bwd_post_vec_kernel(const float* __restrict__ dfhat, const float* __restrict__ f,
                                                               const float* __restrict__ inv_norm,
                                                               const float* __restrict__ bound, float* __restrict__ grad_f,
                                                               int HW, int HWpj, int nslot, int PJ, int NS, int G) {
    constexpr int K = 32 * PER;
    __shared__ float4 s_dot[32][8];
    const int n = blockIdx.y;
    const int q = threadIdx.x & 7, cg = threadIdx.x >> 3;
    const int p = blockIdx.x * 32 + 4 * q;
    const bool live = p < HW;                                   // HW % 4 == 0: a quad is inside or outside as a whole
    const float unscale = 1.f / pcm_grad_scale(bound);
    const size_t split_stride = static_cast<size_t>(K) * HWpj;
    const int nsplit = pcm_unit_segments(pcm_sched(static_cast<int>(gridDim.y) * PJ, NS, G),
                                         n * PJ + (blockIdx.x * 32) / (2 * kTile));
    const float* gsrc = dfhat + static_cast<size_t>(n) * nslot * split_stride + static_cast<size_t>(cg * PER) * HWpj + p;
    const float* fsrc = f + (static_cast<size_t>(n) * K + cg * PER) * HW + p;
    float4 g[PER], fv[PER];
    float4 dot = make_float4(0.f, 0.f, 0.f, 0.f);
    if (live) {
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            fv[j] = __ldg(reinterpret_cast<const float4*>(fsrc + static_cast<size_t>(j) * HW));
            g[j] = __ldg(reinterpret_cast<const float4*>(gsrc + static_cast<size_t>(j) * HWpj));
        }
        if (nsplit == 2) {                                      // the common multi-segment case: everything in flight together
            float4 g2[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) g2[j] = __ldg(reinterpret_cast<const float4*>(gsrc + split_stride + static_cast<size_t>(j) * HWpj));
#pragma unroll
            for (int j = 0; j < PER; ++j) { g[j].x += g2[j].x; g[j].y += g2[j].y; g[j].z += g2[j].z; g[j].w += g2[j].w; }
        } else {
            for (int sp = 1; sp < nsplit; ++sp) {
#pragma unroll
                for (int j = 0; j < PER; ++j) {
                    const float4 t = __ldg(reinterpret_cast<const float4*>(gsrc + sp * split_stride + static_cast<size_t>(j) * HWpj));
                    g[j].x += t.x; g[j].y += t.y; g[j].z += t.z; g[j].w += t.w;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            g[j].x *= unscale; g[j].y *= unscale; g[j].z *= unscale; g[j].w *= unscale;
            dot.x = fmaf(g[j].x, fv[j].x, dot.x);
            dot.y = fmaf(g[j].y, fv[j].y, dot.y);
            dot.z = fmaf(g[j].z, fv[j].z, dot.z);
            dot.w = fmaf(g[j].w, fv[j].w, dot.w);
        }
    }
    s_dot[cg][q] = dot;
    __syncthreads();
    if (!live) return;
    dot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
    for (int i = 0; i < 32; ++i) {                              // fixed order: deterministic
        const float4 t = s_dot[i][q];
        dot.x += t.x; dot.y += t.y; dot.z += t.z; dot.w += t.w;
    }
    const float4 inv = *reinterpret_cast<const float4*>(inv_norm + static_cast<size_t>(n) * HWpj + p);
    auto coef_of = [](float d, float iv) {
        const float nrm = 1.f / iv - 1e-5f;
        return nrm > 0.f ? d * iv * iv / nrm : 0.f;
    };
    const float4 coef = make_float4(coef_of(dot.x, inv.x), coef_of(dot.y, inv.y), coef_of(dot.z, inv.z), coef_of(dot.w, inv.w));
    float* dst = grad_f + (static_cast<size_t>(n) * K + cg * PER) * HW + p;
#pragma unroll
    for (int j = 0; j < PER; ++j)
        *reinterpret_cast<float4*>(dst + static_cast<size_t>(j) * HW) =
            make_float4(g[j].x * inv.x - fv[j].x * coef.x, g[j].y * inv.y - fv[j].y * coef.y,
                        g[j].z * inv.z - fv[j].z * coef.z, g[j].w * inv.w - fv[j].w * coef.w);
}

static size_t fwd_smem_bytes(int K, int wb) {
    const size_t a = static_cast<size_t>(kTile) * K * 2, b = static_cast<size_t>(wb) * K * 2;
    return kABuf * a + kFStages * b + kCamStages * static_cast<size_t>(2 * wb) * 2 * kCamRows + sizeof(FwdBars) + 64;
}

static size_t bwd_smem_bytes(int K, int wb, bool split) {
    const size_t ku = split ? 3 * kAugK : kAugK;
    return static_cast<size_t>(kTile) * ku * 2 + kBwdStages * (static_cast<size_t>(wb) * ku * 2 + static_cast<size_t>(wb) * K * 2) +
           sizeof(BwdBars) + 64;
}

}  // namespace wsegk

using namespace wsegk;

// ---- optional main-kernel timing (wsegk_debug_pcm_profile) ----------------------------------------
namespace {
struct PcmProf {
    bool on = false;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // fwd start/stop, bwd start/stop, fix-up
    bool have[3] = {false, false, false};
};
thread_local PcmProf g_prof;
thread_local long long* g_trace_bwd = nullptr; // the same for the backward kernel (768 int64)
thread_local long long* g_trace = nullptr;   // device buffer of 1024 + 2 * 148 + 2 int64 for the forward kernel's timeline, or null
void prof_mark(int which, wsegk_stream_t stream) {
    if (!g_prof.on) return;
    if (!g_prof.ev[which]) cudaEventCreate(&g_prof.ev[which]);
    cudaEventRecord(g_prof.ev[which], stream);
    if (which & 1) g_prof.have[which >> 1] = true;
}
}  // namespace

extern "C" int wsegk_debug_pcm_profile(int enable) {
    g_prof.on = enable != 0;
    return WSEGK_OK;
}

extern "C" int wsegk_debug_pcm_trace(long long* device_buffer) {
    g_trace = device_buffer;
    return WSEGK_OK;
}

extern "C" int wsegk_debug_pcm_trace_bwd(long long* device_buffer) {
    g_trace_bwd = device_buffer;
    return WSEGK_OK;
}

static int prof_read(int k, float* out) {
    if (!out) return WSEGK_OK;
    *out = -1.f;
    if (!g_prof.have[k]) return WSEGK_OK;
    WSEGK_CUDA_OK(cudaEventSynchronize(g_prof.ev[2 * k + 1]));
    WSEGK_CUDA_OK(cudaEventElapsedTime(out, g_prof.ev[2 * k], g_prof.ev[2 * k + 1]));
    return WSEGK_OK;
}

extern "C" int wsegk_debug_pcm_main_ms(float* fwd_ms, float* bwd_ms) {
    if (int rc = prof_read(0, fwd_ms)) return rc;
    return prof_read(1, bwd_ms);
}

extern "C" int wsegk_debug_pcm_fixup_ms(float* fix_ms) { return prof_read(2, fix_ms); }

static int pcm_check(const char* who, int N, int C, int K, int h, int w) {
    WSEGK_REQUIRE(N > 0 && h > 0 && w > 0, WSEGK_ERR_BAD_SHAPE, "%s: N,h,w must be positive", who);
    WSEGK_REQUIRE(C >= 1 && C <= kMaxClasses, WSEGK_ERR_BAD_SHAPE, "%s: need 1 <= C <= %d (got %d)", who,
                  kMaxClasses, C);
    WSEGK_REQUIRE(K >= 32 && K <= kMaxPcmK && (K % 32) == 0, WSEGK_ERR_BAD_SHAPE,
                  "%s: need K %% 32 == 0 and 32 <= K <= %d (got %d)", who, kMaxPcmK, K);
    WSEGK_REQUIRE(static_cast<long long>(h) * w <= (1 << 22) && static_cast<long long>(N) * h * w < (1ll << 28),
                  WSEGK_ERR_BAD_SHAPE, "%s: problem too large", who);
    return WSEGK_OK;
}

template <int WB, int MODE>
static int launch_fwd(const PcmPlan& pl, int N, wsegk_stream_t stream) {
    const size_t smem = fwd_smem_bytes(pl.K, WB);
    WSEGK_REQUIRE(smem <= 232448, WSEGK_ERR_BAD_SHAPE, "pcm_fwd: %zu bytes of shared memory needed", smem);
    WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_fwd_kernel<WB, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    pcm_fwd_kernel<WB, MODE><<<dim3(2 * pl.G), kFwdThreads, smem, stream>>>(
        reinterpret_cast<const uint8_t*>(pl.fa), reinterpret_cast<const uint8_t*>(pl.fb),
        reinterpret_cast<const uint8_t*>(pl.camb), pl.part, reinterpret_cast<uint2*>(pl.mask),
        reinterpret_cast<uint2*>(pl.nzmask), pl.bound + kBoundLo2, pl.out, pl.dsum, pl.C, pl.K, pl.HW, pl.PJ,
        pl.NS, N * pl.PJ, pl.nsplit, g_trace);
    return WSEGK_OK;
}

template <int WB, bool SPLIT>
static int launch_bwd(const PcmPlan& pl, int N, wsegk_stream_t stream) {
    const size_t smem = bwd_smem_bytes(pl.K, WB, SPLIT);
    WSEGK_REQUIRE(smem <= 232448, WSEGK_ERR_BAD_SHAPE, "pcm_bwd: %zu bytes of shared memory needed", smem);
    WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_bwd_kernel<WB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    pcm_bwd_kernel<WB, SPLIT><<<dim3(2 * pl.Gb), kMainThreads, smem, stream>>>(
        reinterpret_cast<const uint8_t*>(pl.ua), reinterpret_cast<const uint8_t*>(pl.vb),
        reinterpret_cast<const uint8_t*>(pl.fb), reinterpret_cast<const uint4*>(pl.mask), pl.bound, pl.dfhat, pl.K,
        pl.PJ, pl.NS, N * pl.PJ, pl.nsplit, g_trace_bwd);
    return WSEGK_OK;
}

static int dispatch_fwd(const PcmPlan& pl, int N, int mode, wsegk_stream_t stream) {
#define WSEGK_FWD_CASE(WBV)                                            \
    case WBV:                                                          \
        return mode == 2 ? launch_fwd<WBV, 2>(pl, N, stream)           \
               : mode == 1 ? launch_fwd<WBV, 1>(pl, N, stream)         \
                           : launch_fwd<WBV, 0>(pl, N, stream);
    switch (pl.wb) {
        WSEGK_FWD_CASE(64)
        WSEGK_FWD_CASE(56)
        default:
            WSEGK_FWD_CASE(48)
    }
#undef WSEGK_FWD_CASE
}

static int dispatch_bwd(const PcmPlan& pl, int N, bool split, wsegk_stream_t stream) {
#define WSEGK_BWD_CASE(WBV) \
    case WBV:               \
        return split ? launch_bwd<WBV, true>(pl, N, stream) : launch_bwd<WBV, false>(pl, N, stream);
    switch (pl.wb) {
        WSEGK_BWD_CASE(64)
        WSEGK_BWD_CASE(56)
        default:
            WSEGK_BWD_CASE(48)
    }
#undef WSEGK_BWD_CASE
}

// the sign fix-up of the precise mode on a workspace the forward has filled
static int launch_fixup(const PcmPlan& pl, int N, wsegk_stream_t stream) {
    const size_t smem = static_cast<size_t>(kFixTile) * pl.K * 4 + static_cast<size_t>(kFixQueue) * 4;
    const int tiles = pl.HWpj / kFixTile;
    // CTA = (row tile, chunk of stages, image); the chunks below the diagonal exit at once, so about half of the grid
    // works: aim at ~8 working CTAs per SM (four are resident), chunks of at least four stages (the scan reads four per pass)
    const long long cells = static_cast<long long>(tiles) * N * pl.NS;
    int chunk = static_cast<int>(std::min<long long>(std::max<long long>(4, cells / (2 * 8 * 148)), 64));
    {
        static const int forced = [] { const char* e = getenv("WSEGK_FIX_CHUNK"); return e ? atoi(e) : 0; }();   // A/B aid
        if (forced > 0) chunk = forced;
    }
    chunk = std::min(chunk, std::max(1, pl.NS));
    const dim3 grid(tiles, ceil_div(pl.NS, chunk), N);
    prof_mark(4, stream);
#define WSEGK_FIX_CASE(KGV)                                                                                              \
    case KGV:                                                                                                            \
        WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_sign_fixup_kernel<KGV>, cudaFuncAttributeMaxDynamicSharedMemorySize,      \
                                           static_cast<int>(smem)));                                                     \
        pcm_sign_fixup_kernel<KGV><<<grid, kFixThreads, smem, stream>>>(                                                 \
            pl.fh32, reinterpret_cast<const uint4*>(pl.nzmask), pl.mask, pl.HWpj, pl.NS, 2 * pl.wb, chunk);              \
        break;
    switch (pl.K / 16) {           // K is a multiple of 32 (pcm_check)
        WSEGK_FIX_CASE(2)
        WSEGK_FIX_CASE(4)
        WSEGK_FIX_CASE(6)
        WSEGK_FIX_CASE(8)
        WSEGK_FIX_CASE(10)
        default:
            WSEGK_FIX_CASE(12)
    }
#undef WSEGK_FIX_CASE
    prof_mark(5, stream);
    count_launches(1);
    return WSEGK_OK;
}

// mode: 0 inference (no masks), 1 fast (sign mask from the FP16 products), 2 precise (+ near-zero mask and fix-up)
static int pcm_fwd_impl(const float* f, const float* cam_norm, float* cam_rv, int N, int C, int K, int h, int w,
                        void* workspace, size_t workspace_bytes, wsegk_stream_t stream, int mode, bool defer_fixup,
                        float* cam_norm_out) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(f && cam_norm && cam_rv && workspace, WSEGK_ERR_NULL, "pcm_fwd: null pointer");
    if (int rc = pcm_check("pcm_fwd", N, C, K, h, w)) return rc;
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, WSEGK_ERR_BAD_ALIGN,
                  "pcm_fwd: workspace must be 1024-byte aligned");
    PcmPlan pl = make_pcm_plan(workspace, N, C, K, h, w, mode == 2);
    WSEGK_REQUIRE(workspace_bytes >= pl.bytes, WSEGK_ERR_WORKSPACE_TOO_SMALL, "pcm_fwd: workspace %zu < %zu",
                  workspace_bytes, pl.bytes);
    {
        const int tiles = ceil_div(std::max(pl.HWpj, pl.HWpi), kTile);
        uint4 *fa = reinterpret_cast<uint4*>(pl.fa), *fb = reinterpret_cast<uint4*>(pl.fb);
        uint4* cb = reinterpret_cast<uint4*>(pl.camb);
        unsigned int* lo2 = reinterpret_cast<unsigned int*>(pl.bound + kBoundLo2);
        const dim3 grid(tiles + pl.NS, N);
        const bool logits = cam_norm_out != nullptr;
        if (logits) {
            launch_cam_plane_peak(cam_norm, pl.peak, N * C, pl.HW, stream);
            count_launches(1);
        }
        if (mode == 2) {
            WSEGK_CUDA_OK(cudaMemsetAsync(lo2, 0, sizeof(float), stream));
            if (logits)
                pcm_prep_kernel<true, true><<<grid, kPrepThreads, 0, stream>>>(f, fa, fb, pl.fh32, lo2, pl.inv_norm, cam_norm, pl.peak,
                                                                      cam_norm_out, cb, C, K, pl.HW, pl.PJ, pl.NS, pl.wb, tiles);
            else
                pcm_prep_kernel<true, false><<<grid, kPrepThreads, 0, stream>>>(f, fa, fb, pl.fh32, lo2, pl.inv_norm, cam_norm, nullptr,
                                                                       nullptr, cb, C, K, pl.HW, pl.PJ, pl.NS, pl.wb, tiles);
        } else if (logits) {
            pcm_prep_kernel<false, true><<<grid, kPrepThreads, 0, stream>>>(f, fa, fb, pl.fh32, lo2, pl.inv_norm, cam_norm, pl.peak,
                                                                   cam_norm_out, cb, C, K, pl.HW, pl.PJ, pl.NS, pl.wb, tiles);
        } else {
            pcm_prep_kernel<false, false><<<grid, kPrepThreads, 0, stream>>>(f, fa, fb, pl.fh32, lo2, pl.inv_norm, cam_norm, nullptr,
                                                                    nullptr, cb, C, K, pl.HW, pl.PJ, pl.NS, pl.wb, tiles);
        }
    }
    prof_mark(0, stream);
    pl.out = cam_rv;
    if (int rc = dispatch_fwd(pl, N, mode, stream)) return rc;
    prof_mark(1, stream);
#ifdef WSEGK_FWD_NOFUSE
    const int fin_all = 1;
#else
    const int fin_all = 0;
#endif
    if (fin_all || pl.fwd_shared_units) {
        pcm_fwd_finalize_kernel<<<dim3(pl.HWpj / 32, N), 256, 0, stream>>>(pl.part, cam_rv, pl.dsum, C, pl.HW, pl.HWpj,
                                                                             pl.nsplit, pl.PJ, pl.NS, pl.G, fin_all);
        count_launches(1);
    }
    count_launches(2);
    if (mode == 2 && !defer_fixup)
        if (int rc = launch_fixup(pl, N, stream)) return rc;
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_fwd(const float* f, const float* cam_norm, float* cam_rv, int N, int C, int K, int h,
                             int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    return pcm_fwd_impl(f, cam_norm, cam_rv, N, C, K, h, w, workspace, workspace_bytes, stream, pcm_mode() ? 2 : 1, false,
                        nullptr);
}

namespace wsegk {
// operand preparation and finalize pass of several problems in one launch each: block -> (problem, block of the problem)
struct PrepItem {
    const float *f, *cam;
    uint4 *fa, *fb, *camb;
    float *fh32, *inv_norm;
    unsigned int* lo2max;
    int HW, PJ, NS, tiles, bx, block0;      // bx = tiles + NS blocks per image; block0 = first block of the problem
};
struct PrepItems {
    int n, C, K, wb;
    PrepItem it[kMaxFwdItems];
};
__global__ void __launch_bounds__(kPrepThreads) pcm_prep_multi_kernel(const __grid_constant__ PrepItems items) {
    int idx = 0;
    for (int i = 1; i < items.n; ++i)
        if (static_cast<int>(blockIdx.x) >= items.it[i].block0) idx = i;
    const PrepItem& it = items.it[idx];
    const int lb = static_cast<int>(blockIdx.x) - it.block0, n = lb / it.bx, b = lb - n * it.bx;
    if (b < it.tiles)
        pcm_prep_f_body<false>(it.f, it.fa, it.fb, it.fh32, it.lo2max, it.inv_norm, items.K, it.HW, it.PJ, it.NS, items.wb, b, n);
    else
        pcm_prep_cam_body<false>(it.cam, nullptr, nullptr, it.camb, items.C, it.HW, it.NS, items.wb, b - it.tiles, n);
}
struct FinItem {
    const float* part;
    float *out, *dsum;
    int HW, HWpj, nslot, PJ, NS, G, N, bx, block0;   // bx = HWpj / 32 blocks per image
};
struct FinItems {
    int n, C;
    FinItem it[kMaxFwdItems];
};
__global__ void __launch_bounds__(256) pcm_fwd_finalize_multi_kernel(const __grid_constant__ FinItems items) {
    int idx = 0;
    for (int i = 1; i < items.n; ++i)
        if (static_cast<int>(blockIdx.x) >= items.it[i].block0) idx = i;
    const FinItem& it = items.it[idx];
    const int lb = static_cast<int>(blockIdx.x) - it.block0, n = lb / it.bx, b = lb - n * it.bx;
    pcm_fwd_finalize_body(it.part, it.out, it.dsum, items.C, it.HW, it.HWpj, it.nslot, it.PJ, it.NS, it.G, 0, it.N, b, n);
}
}  // namespace wsegk

// ---- several forward-only problems in one launch -----------------------------------------------------
namespace {
constexpr int kMultiWb = 64;         // one kernel instantiation for all problems
constexpr int kMultiPairs = 72;      // inference: the PCM has the GPU to itself
// CTA pairs per problem, in proportion to its stages (at least one; at most half its stages, as for a single problem)
int multi_split(int n_items, const int* N, const int* h, const int* w, int* pairs) {
    long long T[kMaxFwdItems], tot = 0;
    for (int i = 0; i < n_items; ++i) {
        const long long HW = static_cast<long long>(h[i]) * w[i];
        T[i] = N[i] * ((HW + 2 * kTile - 1) / (2 * kTile)) * ((HW + 2 * kMultiWb - 1) / (2 * kMultiWb));
        tot += T[i];
    }
    // every problem starts with one pair; each further pair goes to the problem whose pairs have the most stages to run
    // (greedy min-max: 30 / 380 / 1892 / 5700 stages of the MSF variants of a 320x480 image -> 1 / 4 / 17 / 50 pairs)
    int sum = 0;
    for (int i = 0; i < n_items; ++i) { pairs[i] = 1; ++sum; }
    while (sum < kMultiPairs) {
        int best = -1;
        for (int i = 0; i < n_items; ++i)
            if (pairs[i] < std::max<long long>(1, T[i] / 2) &&
                (best < 0 || T[i] * pairs[best] > T[best] * pairs[i]))
                best = i;
        if (best < 0) break;
        ++pairs[best];
        ++sum;
    }
    return sum;
}
int multi_check(const char* who, int n_items, const int* N, const int* h, const int* w, int C, int K) {
    WSEGK_REQUIRE(N && h && w, WSEGK_ERR_NULL, "%s: null pointer", who);
    WSEGK_REQUIRE(n_items >= 1 && n_items <= kMaxFwdItems, WSEGK_ERR_BAD_SHAPE, "%s: need 1 <= n_items <= %d (got %d)", who,
                  kMaxFwdItems, n_items);
    for (int i = 0; i < n_items; ++i)
        if (int rc = pcm_check(who, N[i], C, K, h[i], w[i])) return rc;
    return WSEGK_OK;
}
}  // namespace

extern "C" int wsegk_pcm_multi_workspace_bytes(int n_items, const int* N, const int* h, const int* w, int C, int K,
                                               size_t* bytes_out) {
    if (int rc = multi_check("pcm_multi_workspace_bytes", n_items, N, h, w, C, K)) return rc;
    WSEGK_REQUIRE(bytes_out, WSEGK_ERR_NULL, "pcm_multi_workspace_bytes: null pointer");
    int pairs[kMaxFwdItems];
    multi_split(n_items, N, h, w, pairs);
    for (int i = 0; i < n_items; ++i) bytes_out[i] = make_pcm_plan(nullptr, N[i], C, K, h[i], w[i], 0, kMultiWb, pairs[i]).bytes;
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_fwd_infer_multi(int n_items, const float* const* f, const float* const* cam_norm,
                                         float* const* cam_rv, const int* N, const int* h, const int* w, int C, int K,
                                         void* const* workspaces, const size_t* workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(f && cam_norm && cam_rv && workspaces && workspace_bytes, WSEGK_ERR_NULL, "pcm_fwd_infer_multi: null pointer");
    if (int rc = multi_check("pcm_fwd_infer_multi", n_items, N, h, w, C, K)) return rc;
    int pairs[kMaxFwdItems];
    const int total_pairs = multi_split(n_items, N, h, w, pairs);
    FwdItems items{};
    items.n = n_items; items.C = C; items.K = K;
    PrepItems prep{};
    prep.n = n_items; prep.C = C; prep.K = K; prep.wb = kMultiWb;
    FinItems fin{};
    fin.C = C;
    int prep_blocks = 0, fin_blocks = 0;
    PcmPlan plans[kMaxFwdItems];
    int g0 = 0;
    for (int i = 0; i < n_items; ++i) {
        WSEGK_REQUIRE(f[i] && cam_norm[i] && cam_rv[i] && workspaces[i], WSEGK_ERR_NULL, "pcm_fwd_infer_multi: null pointer (item %d)", i);
        WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspaces[i]) & 1023) == 0, WSEGK_ERR_BAD_ALIGN,
                      "pcm_fwd_infer_multi: workspace %d must be 1024-byte aligned", i);
        PcmPlan& pl = plans[i];
        pl = make_pcm_plan(workspaces[i], N[i], C, K, h[i], w[i], 0, kMultiWb, pairs[i]);
        WSEGK_REQUIRE(workspace_bytes[i] >= pl.bytes, WSEGK_ERR_WORKSPACE_TOO_SMALL, "pcm_fwd_infer_multi: workspace %d: %zu < %zu",
                      i, workspace_bytes[i], pl.bytes);
        const int tiles = ceil_div(std::max(pl.HWpj, pl.HWpi), kTile);
        PrepItem& pi = prep.it[i];
        pi.f = f[i]; pi.cam = cam_norm[i];
        pi.fa = reinterpret_cast<uint4*>(pl.fa); pi.fb = reinterpret_cast<uint4*>(pl.fb); pi.camb = reinterpret_cast<uint4*>(pl.camb);
        pi.fh32 = pl.fh32; pi.inv_norm = pl.inv_norm; pi.lo2max = reinterpret_cast<unsigned int*>(pl.bound + kBoundLo2);
        pi.HW = pl.HW; pi.PJ = pl.PJ; pi.NS = pl.NS; pi.tiles = tiles; pi.bx = tiles + pl.NS; pi.block0 = prep_blocks;
        prep_blocks += pi.bx * N[i];
        if (pl.fwd_shared_units) {
            FinItem& fi = fin.it[fin.n++];
            fi.part = pl.part; fi.out = cam_rv[i]; fi.dsum = pl.dsum;
            fi.HW = pl.HW; fi.HWpj = pl.HWpj; fi.nslot = pl.nsplit; fi.PJ = pl.PJ; fi.NS = pl.NS; fi.G = pl.G; fi.N = N[i];
            fi.bx = pl.HWpj / 32; fi.block0 = fin_blocks;
            fin_blocks += fi.bx * N[i];
        }
        FwdItem& it = items.it[i];
        it.fa = reinterpret_cast<const uint8_t*>(pl.fa);
        it.fb = reinterpret_cast<const uint8_t*>(pl.fb);
        it.camb = reinterpret_cast<const uint8_t*>(pl.camb);
        it.part = pl.part; it.out = cam_rv[i]; it.dsum = pl.dsum;
        it.HW = pl.HW; it.PJ = pl.PJ; it.NS = pl.NS; it.n_units = N[i] * pl.PJ; it.nslot = pl.nsplit;
        it.g0 = g0; it.G = pairs[i];
        g0 += pairs[i];
    }
    const size_t smem = fwd_smem_bytes(K, kMultiWb);
    WSEGK_REQUIRE(smem <= 232448, WSEGK_ERR_BAD_SHAPE, "pcm_fwd_infer_multi: %zu bytes of shared memory needed", smem);
    WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_fwd_multi_kernel<kMultiWb>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem)));
    pcm_prep_multi_kernel<<<prep_blocks, kPrepThreads, 0, stream>>>(prep);
    prof_mark(0, stream);
    pcm_fwd_multi_kernel<kMultiWb><<<dim3(2 * total_pairs), kFwdThreads, smem, stream>>>(items);
    prof_mark(1, stream);
    int launches = 2;
    if (fin.n > 0) {
        pcm_fwd_finalize_multi_kernel<<<fin_blocks, 256, 0, stream>>>(fin);
        ++launches;
    }
    count_launches(launches);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_fwd_infer(const float* f, const float* cam_norm, float* cam_rv, int N, int C, int K, int h,
                                   int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    return pcm_fwd_impl(f, cam_norm, cam_rv, N, C, K, h, w, workspace, workspace_bytes, stream, 0, false, nullptr);
}

// cam_norm_out != nullptr: `cam_norm` holds raw logits; K1a is fused into the operand preparation
extern "C" int wsegk_pcm_fwd_logits(const float* f, const float* cam_logits, float* cam_norm_out, float* cam_rv, int N, int C,
                                    int K, int h, int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    WSEGK_REQUIRE(cam_norm_out != nullptr, WSEGK_ERR_NULL, "pcm_fwd_logits: null pointer");
    return pcm_fwd_impl(f, cam_logits, cam_rv, N, C, K, h, w, workspace, workspace_bytes, stream, pcm_mode() ? 2 : 1, false,
                        cam_norm_out);
}

extern "C" int wsegk_pcm_fwd_ex(const float* f, const float* cam, float* cam_norm_out, float* cam_rv, int N, int C, int K,
                                int h, int w, void* workspace, size_t workspace_bytes, int flags, wsegk_stream_t stream) {
    WSEGK_REQUIRE((flags & ~(WSEGK_PCM_PRECISE | WSEGK_PCM_DEFER_FIXUP)) == 0, WSEGK_ERR_BAD_SHAPE,
                  "pcm_fwd_ex: unknown flags 0x%x", flags);
    return pcm_fwd_impl(f, cam, cam_rv, N, C, K, h, w, workspace, workspace_bytes, stream,
                        (flags & WSEGK_PCM_PRECISE) ? 2 : 1, (flags & WSEGK_PCM_DEFER_FIXUP) != 0, cam_norm_out);
}

extern "C" int wsegk_pcm_sign_fixup(int N, int C, int K, int h, int w, void* workspace, size_t workspace_bytes,
                                    wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(workspace, WSEGK_ERR_NULL, "pcm_sign_fixup: null pointer");
    if (int rc = pcm_check("pcm_sign_fixup", N, C, K, h, w)) return rc;
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, WSEGK_ERR_BAD_ALIGN,
                  "pcm_sign_fixup: workspace must be 1024-byte aligned");
    PcmPlan pl = make_pcm_plan(workspace, N, C, K, h, w, 1);
    WSEGK_REQUIRE(workspace_bytes >= pl.bytes, WSEGK_ERR_WORKSPACE_TOO_SMALL, "pcm_sign_fixup: workspace %zu < %zu",
                  workspace_bytes, pl.bytes);
    if (int rc = launch_fixup(pl, N, stream)) return rc;
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

static int pcm_bwd_impl(const float* f, const float* cam_norm, const float* cam_rv, const float* grad_cam_rv,
                        float* grad_f, int N, int C, int K, int h, int w, void* workspace, size_t workspace_bytes,
                        int precise, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(f && cam_norm && cam_rv && grad_cam_rv && grad_f && workspace, WSEGK_ERR_NULL,
                  "pcm_bwd: null pointer");
    if (int rc = pcm_check("pcm_bwd", N, C, K, h, w)) return rc;
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, WSEGK_ERR_BAD_ALIGN,
                  "pcm_bwd: workspace must be 1024-byte aligned");
    PcmPlan pl = make_pcm_plan(workspace, N, C, K, h, w, precise);
    WSEGK_REQUIRE(workspace_bytes >= pl.bytes, WSEGK_ERR_WORKSPACE_TOO_SMALL, "pcm_bwd: workspace %zu < %zu",
                  workspace_bytes, pl.bytes);
    WSEGK_CUDA_OK(cudaMemsetAsync(pl.bound, 0, sizeof(float), stream));
    {
        const int tiles = ceil_div(std::max(pl.HWpj, pl.HWpi), kTile);
        if (precise)
            pcm_bwd_prep_kernel<true><<<dim3(tiles, N), 128, 0, stream>>>(
                cam_norm, cam_rv, grad_cam_rv, pl.dsum, reinterpret_cast<uint4*>(pl.ua), reinterpret_cast<uint4*>(pl.vb),
                reinterpret_cast<unsigned int*>(pl.bound), C, pl.HW, pl.PJ, pl.NS, pl.wb);
        else
            pcm_bwd_prep_kernel<false><<<dim3(tiles, N), 128, 0, stream>>>(
                cam_norm, cam_rv, grad_cam_rv, pl.dsum, reinterpret_cast<uint4*>(pl.ua), reinterpret_cast<uint4*>(pl.vb),
                reinterpret_cast<unsigned int*>(pl.bound), C, pl.HW, pl.PJ, pl.NS, pl.wb);
    }
    prof_mark(2, stream);
    if (int rc = dispatch_bwd(pl, N, precise != 0, stream)) return rc;
    prof_mark(3, stream);
    if ((pl.HW & 3) == 0 && (reinterpret_cast<uintptr_t>(f) & 15) == 0 && (reinterpret_cast<uintptr_t>(grad_f) & 15) == 0) {
        const dim3 grid(ceil_div(pl.HW, 32), N);
#define WSEGK_POST_CASE(PERV)                                                                                   \
    case PERV:                                                                                                  \
        pcm_bwd_post_vec_kernel<PERV><<<grid, 256, 0, stream>>>(pl.dfhat, f, pl.inv_norm, pl.bound, grad_f, pl.HW, \
                                                                pl.HWpj, pl.nsplit, pl.PJ, pl.NS, pl.Gb);        \
        break;
        switch (K / 32) {
            WSEGK_POST_CASE(1)
            WSEGK_POST_CASE(2)
            WSEGK_POST_CASE(3)
            WSEGK_POST_CASE(4)
            WSEGK_POST_CASE(5)
            default:
                WSEGK_POST_CASE(6)
        }
#undef WSEGK_POST_CASE
    } else {
        pcm_bwd_post_kernel<<<dim3(ceil_div(pl.HW, kPostPix), N), 256, 0, stream>>>(
            pl.dfhat, f, pl.inv_norm, pl.bound, grad_f, K, pl.HW, pl.HWpj, pl.nsplit, pl.PJ, pl.NS, pl.Gb);
    }
    count_launches(3);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_bwd(const float* f, const float* cam_norm, const float* cam_rv, const float* grad_cam_rv,
                             float* grad_f, int N, int C, int K, int h, int w, void* workspace,
                             size_t workspace_bytes, wsegk_stream_t stream) {
    return pcm_bwd_impl(f, cam_norm, cam_rv, grad_cam_rv, grad_f, N, C, K, h, w, workspace, workspace_bytes, pcm_mode(),
                        stream);
}

extern "C" int wsegk_pcm_bwd_ex(const float* f, const float* cam_norm, const float* cam_rv, const float* grad_cam_rv,
                                float* grad_f, int N, int C, int K, int h, int w, void* workspace, size_t workspace_bytes,
                                int flags, wsegk_stream_t stream) {
    WSEGK_REQUIRE((flags & ~(WSEGK_PCM_PRECISE | WSEGK_PCM_DEFER_FIXUP)) == 0, WSEGK_ERR_BAD_SHAPE,
                  "pcm_bwd_ex: unknown flags 0x%x", flags);
    return pcm_bwd_impl(f, cam_norm, cam_rv, grad_cam_rv, grad_f, N, C, K, h, w, workspace, workspace_bytes,
                        (flags & WSEGK_PCM_PRECISE) ? 1 : 0, stream);
}
