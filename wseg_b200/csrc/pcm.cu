// K2 — SEAM pixel correlation module (replaces network/resnet38_contrast.py:70-73 after the f9 conv).
//
//   fh_i = f_i / (|f_i| + 1e-5)              (per pixel, over K channels)
//   S    = fh fh^T  (HW x HW, symmetric),  A = relu(S),  D_j = sum_i A_ij + 1e-5
//   out[c,j] = sum_i cam[c,i] A_ij / D_j
//
// Design (B200): the HWxHW affinity never exists in HBM.  A CTA owns a 128-pixel j-tile and streams
// 128-pixel i-tiles through shared memory with bulk TMA copies; tcgen05 computes S^T tile = Fj Fi^T into
// TMEM, four warps apply ReLU and write the BF16 result back into TMEM, and a second tcgen05 MMA with the
// A operand in TMEM multiplies it by the [cam_hi ; 1 ; cam_lo] rows of the same i-tile, so numerators and
// the column sum D_j come out of one accumulator.  Operands are stored by the prep kernels directly in
// the UMMA canonical no-swizzle layout, so every tile is one contiguous bulk copy (no tensor maps).
#include <cuda_bf16.h>

#include "common.cuh"
#include "ptx.cuh"

namespace wsegk {

// ------------------------------------------------------------------------------------------------
// operand preparation
// ------------------------------------------------------------------------------------------------
// Canonical K-major tile of R rows x KK bf16 columns: 16-byte chunk (row r, column group g) lives at
// chunk index g*R + r, i.e. core matrices (8 rows x 16 B) are 128 contiguous bytes, SBO = 128 B,
// LBO = R*16 B.

// f [N,K,HW] fp32 -> fhat tiles (bf16) + inv_norm.  One CTA per (tile, image); 256 threads.
__global__ void __launch_bounds__(256) pcm_prep_f_kernel(const float* __restrict__ f,
                                                         uint16_t* __restrict__ fhat,
                                                         float* __restrict__ inv_norm, int K, int HW,
                                                         int T, int HWp) {
    extern __shared__ float s_val[];            // [K][129]
    __shared__ float s_part[2][kTile];
    __shared__ float s_inv[kTile];
    const int t = blockIdx.x, n = blockIdx.y;
    const int px = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int p = t * kTile + px;
    const float* src = f + static_cast<size_t>(n) * K * HW + p;
    float ss = 0.f;
    for (int k = half; k < K; k += 2) {
        float v = (p < HW) ? __ldg(src + static_cast<size_t>(k) * HW) : 0.f;
        s_val[k * 129 + px] = v;
        ss = fmaf(v, v, ss);
    }
    s_part[half][px] = ss;
    __syncthreads();
    if (threadIdx.x < kTile) {
        float nrm = sqrtf(s_part[0][px] + s_part[1][px]);
        float inv = 1.f / (nrm + 1e-5f);
        s_inv[px] = inv;
        inv_norm[static_cast<size_t>(n) * HWp + p] = (p < HW) ? inv : 0.f;
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(fhat + (static_cast<size_t>(n) * T + t) * kTile * K);
    const int groups = K >> 3;
    for (int chunk = threadIdx.x; chunk < groups * kTile; chunk += blockDim.x) {
        const int g = chunk >> 7, r = chunk & 127;
        const float inv = s_inv[r];
        const float* v = s_val + (g * 8) * 129 + r;
        uint4 o;
        o.x = pack_bf16x2(v[0 * 129] * inv, v[1 * 129] * inv);
        o.y = pack_bf16x2(v[2 * 129] * inv, v[3 * 129] * inv);
        o.z = pack_bf16x2(v[4 * 129] * inv, v[5 * 129] * inv);
        o.w = pack_bf16x2(v[6 * 129] * inv, v[7 * 129] * inv);
        dst[chunk] = o;
    }
}

// cam_norm [N,C,HW] fp32 -> B-operand tiles [48 rows x 128 pixels] bf16, K-major (K = pixel):
//   rows 0..C-1 = bf16(cam), row C = 1 (column-sum row), rows 24..24+C-1 = bf16(cam - hi), others 0.
__global__ void __launch_bounds__(128) pcm_prep_cam_kernel(const float* __restrict__ cam,
                                                           uint16_t* __restrict__ camb, int C, int HW,
                                                           int T) {
    __shared__ __align__(16) uint16_t s_tile[kCamRows * kTile];   // chunk (row, pixel group g) at g*48 + row
    const int t = blockIdx.x, n = blockIdx.y;
    const int px = threadIdx.x;
    const int p = t * kTile + px;
    const bool live = p < HW;
    const float* src = cam + static_cast<size_t>(n) * C * HW + p;
    const int g = px >> 3, e = px & 7;
    auto put = [&](int row, float v) {
        s_tile[(g * kCamRows + row) * 8 + e] = __bfloat16_as_ushort(__float2bfloat16_rn(v));
    };
    for (int c = 0; c < kCamHalf; ++c) {
        float hi = 0.f, lo = 0.f;
        if (live && c < C) {
            float v = __ldg(src + static_cast<size_t>(c) * HW);
            hi = __bfloat162float(__float2bfloat16_rn(v));
            lo = v - hi;
        } else if (live && c == C) {
            hi = 1.f;
        }
        put(c, hi);
        put(kCamHalf + c, lo);
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(camb + (static_cast<size_t>(n) * T + t) * kCamRows * kTile);
    const uint4* s4 = reinterpret_cast<const uint4*>(s_tile);
    for (int chunk = threadIdx.x; chunk < kCamRows * (kTile / 8); chunk += blockDim.x) dst[chunk] = s4[chunk];
}

// ------------------------------------------------------------------------------------------------
// forward main kernel
// ------------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 192;   // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: ReLU/convert + epilogue
constexpr int kStages = 2;
constexpr uint32_t kCamTileBytes = kCamRows * kTile * 2;   // 12288
// TMEM columns
constexpr uint32_t kColS = 0;      // 2 x 128 fp32 columns: S^T tile, double buffered
constexpr uint32_t kColP = 256;    // 2 x 64 columns: relu(S) as packed bf16 (A operand of the 2nd MMA)
constexpr uint32_t kColO = 384;    // 48 fp32 columns: [num_hi | colsum | num_lo] accumulator
constexpr uint32_t kTmemCols = 512;

struct FwdBars {
    uint64_t fj, full[kStages], empty[kStages], s_full[2], s_free[2], p_full[2], p_free[2], o_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kFwdThreads, 1)
pcm_fwd_kernel(const uint16_t* __restrict__ fhat, const uint16_t* __restrict__ camb,
               float* __restrict__ part, int K, int T, int HWp, int nsplit) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tile_bytes = kTile * K * 2;
    uint8_t* s_fj = smem;
    uint8_t* s_stage = smem + tile_bytes;                         // kStages x (Fi tile + cam tile)
    const uint32_t stage_bytes = tile_bytes + kCamTileBytes;
    FwdBars* bars = reinterpret_cast<FwdBars*>(s_stage + kStages * stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tj = blockIdx.x, sp = blockIdx.y, n = blockIdx.z;
    const int i_begin = (T * sp) / nsplit, i_end = (T * (sp + 1)) / nsplit;
    const int ntiles = i_end - i_begin;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars->fj), 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&bars->s_full[b]), 1);
            mbar_init(smem_u32(&bars->s_free[b]), 4);
            mbar_init(smem_u32(&bars->p_full[b]), 4);
            mbar_init(smem_u32(&bars->p_free[b]), 1);
        }
        mbar_init(smem_u32(&bars->o_full), 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    const uint8_t* g_f = reinterpret_cast<const uint8_t*>(fhat) + static_cast<size_t>(n) * T * tile_bytes;
    const uint8_t* g_c = reinterpret_cast<const uint8_t*>(camb) + static_cast<size_t>(n) * T * kCamTileBytes;

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA producer
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(&bars->fj), tile_bytes);
            tma_load_1d(smem_u32(s_fj), g_f + static_cast<size_t>(tj) * tile_bytes, tile_bytes,
                        smem_u32(&bars->fj));
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages, u = t / kStages;
                mbar_wait(smem_u32(&bars->empty[s]), (u & 1) ^ 1);
                const uint32_t bar = smem_u32(&bars->full[s]);
                uint8_t* dst = s_stage + s * stage_bytes;
                mbar_arrive_expect_tx(bar, stage_bytes);
                tma_load_1d(smem_u32(dst), g_f + static_cast<size_t>(i_begin + t) * tile_bytes, tile_bytes, bar);
                tma_load_1d(smem_u32(dst + tile_bytes), g_c + static_cast<size_t>(i_begin + t) * kCamTileBytes,
                            kCamTileBytes, bar);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc1 = umma_idesc_bf16(kTile, kTile, 0, 0);
            const uint32_t idesc2 = umma_idesc_bf16(kTile, kCamRows, 0, 0);
            const uint32_t lbo_f = kTile * 16;        // bytes between the two K halves of an MMA (16 row groups)
            const uint32_t lbo_c = kCamRows * 16;
            const int ksteps = K >> 4;
            mbar_wait(smem_u32(&bars->fj), 0);
            auto issue_s = [&](int t) {
                const int s = t % kStages, u = t / kStages, b = t & 1, ub = t >> 1;
                mbar_wait(smem_u32(&bars->full[s]), u & 1);
                mbar_wait(smem_u32(&bars->s_free[b]), (ub & 1) ^ 1);
                tc_fence_after();
                const uint32_t a0 = smem_u32(s_fj), b0 = smem_u32(s_stage + s * stage_bytes);
                for (int ks = 0; ks < ksteps; ++ks) {
                    umma_ss(tmem + kColS + b * 128, umma_smem_desc(a0 + ks * 2 * lbo_f, lbo_f, 128),
                            umma_smem_desc(b0 + ks * 2 * lbo_f, lbo_f, 128), idesc1, ks > 0);
                }
                umma_commit(smem_u32(&bars->s_full[b]));
            };
            issue_s(0);
            for (int t = 0; t < ntiles; ++t) {
                if (t + 1 < ntiles) issue_s(t + 1);
                const int s = t % kStages, b = t & 1, ub = t >> 1;
                mbar_wait(smem_u32(&bars->p_full[b]), ub & 1);
                tc_fence_after();
                const uint32_t c0 = smem_u32(s_stage + s * stage_bytes + tile_bytes);
                for (int ks = 0; ks < kTile / 16; ++ks) {
                    umma_ts(tmem + kColO, tmem + kColP + b * 64 + ks * 8,
                            umma_smem_desc(c0 + ks * 2 * lbo_c, lbo_c, 128), idesc2, (t > 0) || (ks > 0));
                }
                umma_commit(smem_u32(&bars->empty[s]));
                umma_commit(smem_u32(&bars->p_free[b]));
            }
            umma_commit(smem_u32(&bars->o_full));
        }
    } else {
        // ---------------------------------------------------------------- ReLU / convert warps
        const int q = warp & 3;                         // TMEM lane quarter this warp may touch
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        for (int t = 0; t < ntiles; ++t) {
            const int b = t & 1, ub = t >> 1;
            mbar_wait(smem_u32(&bars->s_full[b]), ub & 1);
            mbar_wait(smem_u32(&bars->p_free[b]), (ub & 1) ^ 1);
            tc_fence_after();
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t r[32], pk[16];
                tmem_ld32(tmem + lane_addr + kColS + b * 128 + c4 * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int m = 0; m < 16; ++m)
                    pk[m] = pack_bf16x2_relu(__uint_as_float(r[2 * m]), __uint_as_float(r[2 * m + 1]));
                tmem_st16(tmem + lane_addr + kColP + b * 64 + c4 * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(smem_u32(&bars->s_free[b]));
                mbar_arrive(smem_u32(&bars->p_full[b]));
            }
        }
        // epilogue: accumulator rows -> partial buffer [n][sp][HWp][48]
        mbar_wait(smem_u32(&bars->o_full), 0);
        tc_fence_after();
        const int row = tj * kTile + q * 32 + lane;
        float4* dst = reinterpret_cast<float4*>(
            part + ((static_cast<size_t>(n) * nsplit + sp) * HWp + row) * kCamRows);
#pragma unroll
        for (int c = 0; c < kCamRows / 16; ++c) {
            uint32_t r[16];
            tmem_ld16(tmem + lane_addr + kColO + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < 4; ++v)
                dst[c * 4 + v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                             __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// out[n,c,j] = (sum_sp hi[c] + lo[c]) / (sum_sp colsum + 1e-5)
__global__ void __launch_bounds__(128) pcm_fwd_finalize_kernel(const float* __restrict__ part,
                                                               float* __restrict__ out,
                                                               float* __restrict__ dsum, int C, int HW,
                                                               int HWp, int nsplit) {
    const int n = blockIdx.y;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= HWp) return;
    float acc[kCamRows];
#pragma unroll
    for (int c = 0; c < kCamRows; ++c) acc[c] = 0.f;
    for (int sp = 0; sp < nsplit; ++sp) {
        const float4* src = reinterpret_cast<const float4*>(
            part + ((static_cast<size_t>(n) * nsplit + sp) * HWp + j) * kCamRows);
#pragma unroll
        for (int v = 0; v < kCamRows / 4; ++v) {
            float4 x = __ldg(src + v);
            acc[4 * v] += x.x; acc[4 * v + 1] += x.y; acc[4 * v + 2] += x.z; acc[4 * v + 3] += x.w;
        }
    }
    float d = 0.f;
#pragma unroll
    for (int c = 0; c < kCamHalf; ++c) if (c == C) d = acc[c];
    d += 1e-5f;
    dsum[static_cast<size_t>(n) * HWp + j] = d;
    if (j >= HW) return;
    float* dst = out + static_cast<size_t>(n) * C * HW + j;
#pragma unroll
    for (int c = 0; c < kCamHalf; ++c)
        if (c < C) dst[static_cast<size_t>(c) * HW] = (acc[c] + acc[kCamHalf + c]) / d;
}


// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// With G = dL/d(out), g_j = G_j / D_j, delta_j = sum_c out[c,j] G[c,j], d_j = delta_j / D_j:
//   dL/dA_ij = cam_i . g_j - d_j,   dL/dS = dL/dA * [S > 0],   dL/dfh = (dS + dS^T) fh
// and (dS + dS^T)_ji = [S_ji > 0] * (U_j . V_i) with the rank-(2C+2) factors
//   U_j = [ g_j ; cam_j ; -d_j ; 1 ],   V_i = [ cam_i ; g_i ; 1 ; -d_i ]              
// so a CTA owning a j-tile needs, per i-tile: S = Fj Fi^T (K), Tt = Uj Vi^T (48), the masked product E in
// TMEM, and dFh_j += E Fi (Fi read MN-major from the same shared-memory tile).  No atomics.

// U/V tiles: [128 pixels x 48] bf16, canonical K-major (chunk (pixel r, group g) at g*128 + r).
__global__ void __launch_bounds__(128) pcm_bwd_prep_kernel(const float* __restrict__ cam,
                                                           const float* __restrict__ out,
                                                           const float* __restrict__ gout,
                                                           const float* __restrict__ dsum,
                                                           uint16_t* __restrict__ uaug,
                                                           uint16_t* __restrict__ vaug, int C, int HW,
                                                           int T, int HWp) {
    __shared__ float s_u[kAugRows][kTile];
    __shared__ float s_v[kAugRows][kTile];
    const int t = blockIdx.x, n = blockIdx.y;
    const int px = threadIdx.x;
    const int p = t * kTile + px;
    for (int i = 0; i < kAugRows; ++i) { s_u[i][px] = 0.f; s_v[i][px] = 0.f; }
    if (p < HW) {
        const float invd = 1.f / dsum[static_cast<size_t>(n) * HWp + p];
        const size_t base = static_cast<size_t>(n) * C * HW + p;
        float delta = 0.f;
        for (int c = 0; c < C; ++c) {
            const float cm = __ldg(cam + base + static_cast<size_t>(c) * HW);
            const float g = __ldg(gout + base + static_cast<size_t>(c) * HW);
            const float o = __ldg(out + base + static_cast<size_t>(c) * HW);
            delta = fmaf(o, g, delta);
            const float gs = g * invd;
            s_u[c][px] = gs;      s_v[c][px] = cm;
            s_u[C + c][px] = cm;  s_v[C + c][px] = gs;
        }
        const float dj = delta * invd;
        s_u[2 * C][px] = -dj;     s_v[2 * C][px] = 1.f;
        s_u[2 * C + 1][px] = 1.f; s_v[2 * C + 1][px] = -dj;
    }
    // each thread only reads back its own column: no barrier needed
    uint4* du = reinterpret_cast<uint4*>(uaug + (static_cast<size_t>(n) * T + t) * kTile * kAugRows);
    uint4* dv = reinterpret_cast<uint4*>(vaug + (static_cast<size_t>(n) * T + t) * kTile * kAugRows);
#pragma unroll
    for (int g = 0; g < kAugRows / 8; ++g) {
        uint4 a, b;
        a.x = pack_bf16x2(s_u[8 * g + 0][px], s_u[8 * g + 1][px]); a.y = pack_bf16x2(s_u[8 * g + 2][px], s_u[8 * g + 3][px]);
        a.z = pack_bf16x2(s_u[8 * g + 4][px], s_u[8 * g + 5][px]); a.w = pack_bf16x2(s_u[8 * g + 6][px], s_u[8 * g + 7][px]);
        b.x = pack_bf16x2(s_v[8 * g + 0][px], s_v[8 * g + 1][px]); b.y = pack_bf16x2(s_v[8 * g + 2][px], s_v[8 * g + 3][px]);
        b.z = pack_bf16x2(s_v[8 * g + 4][px], s_v[8 * g + 5][px]); b.w = pack_bf16x2(s_v[8 * g + 6][px], s_v[8 * g + 7][px]);
        du[g * kTile + px] = a;
        dv[g * kTile + px] = b;
    }
}

constexpr uint32_t kAugTileBytes = kTile * kAugRows * 2;   // 12288
// TMEM columns (backward)
constexpr uint32_t kColBS = 0;      // 128: S tile
constexpr uint32_t kColBT = 128;    // 128: Tt = U V^T tile
constexpr uint32_t kColBG = 256;    // K (<=192): dFh accumulator
constexpr uint32_t kColBE = 448;    // 64: masked product as packed bf16

struct BwdBars {
    uint64_t fj, full[kStages], empty[kStages], st_full, e_full, g_full;
    uint32_t tmem_base;
};

__global__ void __launch_bounds__(kFwdThreads, 1)
pcm_bwd_kernel(const uint16_t* __restrict__ fhat, const uint16_t* __restrict__ uaug,
               const uint16_t* __restrict__ vaug, float* __restrict__ dfhat, int K, int T, int HWp,
               int nsplit) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t tile_bytes = kTile * K * 2;
    uint8_t* s_fj = smem;
    uint8_t* s_uj = smem + tile_bytes;
    uint8_t* s_stage = s_uj + kAugTileBytes;
    const uint32_t stage_bytes = tile_bytes + kAugTileBytes;
    BwdBars* bars = reinterpret_cast<BwdBars*>(s_stage + kStages * stage_bytes);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tj = blockIdx.x, sp = blockIdx.y, n = blockIdx.z;
    const int i_begin = (T * sp) / nsplit, i_end = (T * (sp + 1)) / nsplit;
    const int ntiles = i_end - i_begin;

    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bars->fj), 1);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(smem_u32(&bars->full[s]), 1);
            mbar_init(smem_u32(&bars->empty[s]), 1);
        }
        mbar_init(smem_u32(&bars->st_full), 1);
        mbar_init(smem_u32(&bars->e_full), 4);
        mbar_init(smem_u32(&bars->g_full), 1);
        mbar_fence_init();
    }
    if (warp == 1) {
        tmem_alloc(smem_u32(&bars->tmem_base), kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = bars->tmem_base;

    const uint8_t* g_f = reinterpret_cast<const uint8_t*>(fhat) + static_cast<size_t>(n) * T * tile_bytes;
    const uint8_t* g_u = reinterpret_cast<const uint8_t*>(uaug) + static_cast<size_t>(n) * T * kAugTileBytes;
    const uint8_t* g_v = reinterpret_cast<const uint8_t*>(vaug) + static_cast<size_t>(n) * T * kAugTileBytes;

    if (warp == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(smem_u32(&bars->fj), tile_bytes + kAugTileBytes);
            tma_load_1d(smem_u32(s_fj), g_f + static_cast<size_t>(tj) * tile_bytes, tile_bytes, smem_u32(&bars->fj));
            tma_load_1d(smem_u32(s_uj), g_u + static_cast<size_t>(tj) * kAugTileBytes, kAugTileBytes,
                        smem_u32(&bars->fj));
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages, u = t / kStages;
                mbar_wait(smem_u32(&bars->empty[s]), (u & 1) ^ 1);
                const uint32_t bar = smem_u32(&bars->full[s]);
                uint8_t* dst = s_stage + s * stage_bytes;
                mbar_arrive_expect_tx(bar, stage_bytes);
                tma_load_1d(smem_u32(dst), g_f + static_cast<size_t>(i_begin + t) * tile_bytes, tile_bytes, bar);
                tma_load_1d(smem_u32(dst + tile_bytes), g_v + static_cast<size_t>(i_begin + t) * kAugTileBytes,
                            kAugTileBytes, bar);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc_s = umma_idesc_bf16(kTile, kTile, 0, 0);
            const uint32_t idesc_g = umma_idesc_bf16(kTile, K, 0, 1);     // B = Fi read MN-major (N = channel)
            const uint32_t lbo = kTile * 16;                              // 2048: K-direction step of a 128-row tile
            mbar_wait(smem_u32(&bars->fj), 0);
            for (int t = 0; t < ntiles; ++t) {
                const int s = t % kStages, u = t / kStages;
                mbar_wait(smem_u32(&bars->full[s]), u & 1);
                tc_fence_after();
                const uint32_t fj0 = smem_u32(s_fj), uj0 = smem_u32(s_uj);
                const uint32_t fi0 = smem_u32(s_stage + s * stage_bytes), vi0 = fi0 + tile_bytes;
                // (MMAs execute in issue order, so these overwrite S/Tt only after the previous E was consumed)
                for (int ks = 0; ks < (K >> 4); ++ks)
                    umma_ss(tmem + kColBS, umma_smem_desc(fj0 + ks * 2 * lbo, lbo, 128),
                            umma_smem_desc(fi0 + ks * 2 * lbo, lbo, 128), idesc_s, ks > 0);
                for (int ks = 0; ks < kAugRows / 16; ++ks)
                    umma_ss(tmem + kColBT, umma_smem_desc(uj0 + ks * 2 * lbo, lbo, 128),
                            umma_smem_desc(vi0 + ks * 2 * lbo, lbo, 128), idesc_s, ks > 0);
                umma_commit(smem_u32(&bars->st_full));
                mbar_wait(smem_u32(&bars->e_full), t & 1);
                tc_fence_after();
                // dFh[j, ch] += E[j, i] * Fi[i, ch]: B is N(ch) x K(pixel), MN-major: LBO = pixel-group step,
                // SBO = channel-group step
                for (int ks = 0; ks < kTile / 16; ++ks)
                    umma_ts(tmem + kColBG, tmem + kColBE + ks * 8, umma_smem_desc(fi0 + ks * 256, 128, lbo),
                            idesc_g, (t > 0) || (ks > 0));
                umma_commit(smem_u32(&bars->empty[s]));
            }
            umma_commit(smem_u32(&bars->g_full));
        }
    } else {
        const int q = warp & 3;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        for (int t = 0; t < ntiles; ++t) {
            mbar_wait(smem_u32(&bars->st_full), t & 1);
            tc_fence_after();
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
                uint32_t sv[32], tv[32], pk[16];
                tmem_ld32(tmem + lane_addr + kColBS + c4 * 32, sv);
                tmem_ld32(tmem + lane_addr + kColBT + c4 * 32, tv);
                tmem_ld_wait();
#pragma unroll
                for (int m = 0; m < 16; ++m) {
                    const float a = __uint_as_float(sv[2 * m]) > 0.f ? __uint_as_float(tv[2 * m]) : 0.f;
                    const float b = __uint_as_float(sv[2 * m + 1]) > 0.f ? __uint_as_float(tv[2 * m + 1]) : 0.f;
                    pk[m] = pack_bf16x2(a, b);
                }
                tmem_st16(tmem + lane_addr + kColBE + c4 * 16, pk);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&bars->e_full));
        }
        mbar_wait(smem_u32(&bars->g_full), 0);
        tc_fence_after();
        const int row = tj * kTile + q * 32 + lane;
        float4* dst = reinterpret_cast<float4*>(dfhat + ((static_cast<size_t>(n) * nsplit + sp) * HWp + row) * K);
        for (int c = 0; c < (K >> 4); ++c) {
            uint32_t r[16];
            tmem_ld16(tmem + lane_addr + kColBG + c * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int v = 0; v < 4; ++v)
                dst[c * 4 + v] = make_float4(__uint_as_float(r[4 * v]), __uint_as_float(r[4 * v + 1]),
                                             __uint_as_float(r[4 * v + 2]), __uint_as_float(r[4 * v + 3]));
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, kTmemCols);
}

// grad_f[n,k,p] = dfh[k] * inv - f[k] * (sum_k dfh[k] f[k]) * inv^2 / |f|,  inv = 1/(|f| + 1e-5)
// (autograd of f / (norm(f) + 1e-5), network/resnet38_contrast.py:70); dfh summed over the i-splits.
__global__ void __launch_bounds__(256) pcm_bwd_post_kernel(const float* __restrict__ dfhat,
                                                           const float* __restrict__ f,
                                                           const float* __restrict__ inv_norm,
                                                           float* __restrict__ grad_f, int K, int HW,
                                                           int HWp, int nsplit) {
    extern __shared__ float s_g[];        // [K][129]
    __shared__ float s_dot[2][kTile];
    const int t = blockIdx.x, n = blockIdx.y;
    // coalesced read of the pixel-major gradient rows
    for (int idx = threadIdx.x; idx < kTile * K; idx += blockDim.x) {
        const int r = idx / K, k = idx - r * K;
        float acc = 0.f;
        for (int sp = 0; sp < nsplit; ++sp)
            acc += __ldg(dfhat + ((static_cast<size_t>(n) * nsplit + sp) * HWp + t * kTile + r) * K + k);
        s_g[k * 129 + r] = acc;
    }
    __syncthreads();
    const int px = threadIdx.x & 127, half = threadIdx.x >> 7;
    const int p = t * kTile + px;
    const bool live = p < HW;
    const float* fp = f + static_cast<size_t>(n) * K * HW + p;
    float dot = 0.f;
    if (live)
        for (int k = half; k < K; k += 2) dot = fmaf(s_g[k * 129 + px], __ldg(fp + static_cast<size_t>(k) * HW), dot);
    s_dot[half][px] = dot;
    __syncthreads();
    if (!live) return;
    dot = s_dot[0][px] + s_dot[1][px];
    const float inv = inv_norm[static_cast<size_t>(n) * HWp + p];
    const float nrm = 1.f / inv - 1e-5f;
    const float coef = nrm > 0.f ? dot * inv * inv / nrm : 0.f;
    float* gp = grad_f + static_cast<size_t>(n) * K * HW + p;
    for (int k = half; k < K; k += 2)
        gp[static_cast<size_t>(k) * HW] = s_g[k * 129 + px] * inv - __ldg(fp + static_cast<size_t>(k) * HW) * coef;
}

static size_t bwd_smem_bytes(int K) {
    return static_cast<size_t>(kTile) * K * 2 * (1 + kStages) + (1 + kStages) * kAugTileBytes + sizeof(BwdBars) + 64;
}

static size_t fwd_smem_bytes(int K) {
    return static_cast<size_t>(kTile) * K * 2 * (1 + kStages) + kStages * kCamTileBytes + sizeof(FwdBars) + 64;
}

}  // namespace wsegk

using namespace wsegk;

// ---- optional main-kernel timing (wsegk_debug_pcm_profile) ----------------------------------------
namespace {
struct PcmProf {
    bool on = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // fwd start/stop, bwd start/stop
    bool have[2] = {false, false};
};
thread_local PcmProf g_prof;
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

extern "C" int wsegk_debug_pcm_main_ms(float* fwd_ms, float* bwd_ms) {
    float* out[2] = {fwd_ms, bwd_ms};
    for (int k = 0; k < 2; ++k) {
        if (!out[k]) continue;
        *out[k] = -1.f;
        if (!g_prof.have[k]) continue;
        WSEGK_CUDA_OK(cudaEventSynchronize(g_prof.ev[2 * k + 1]));
        WSEGK_CUDA_OK(cudaEventElapsedTime(out[k], g_prof.ev[2 * k], g_prof.ev[2 * k + 1]));
    }
    return WSEGK_OK;
}

static int pcm_check(const char* who, int N, int C, int K, int h, int w) {
    WSEGK_REQUIRE(N > 0 && h > 0 && w > 0, WSEGK_ERR_BAD_SHAPE, "%s: N,h,w must be positive", who);
    WSEGK_REQUIRE(C >= 1 && C <= kMaxClasses, WSEGK_ERR_BAD_SHAPE, "%s: need 1 <= C <= %d (got %d)", who,
                  kMaxClasses, C);
    WSEGK_REQUIRE(K >= 16 && K <= kMaxPcmK && (K % 16) == 0, WSEGK_ERR_BAD_SHAPE,
                  "%s: need K %% 16 == 0 and 16 <= K <= %d (got %d)", who, kMaxPcmK, K);
    WSEGK_REQUIRE(static_cast<long long>(N) * ceil_div(h * w, kTile) < (1ll << 30), WSEGK_ERR_BAD_SHAPE,
                  "%s: problem too large", who);
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_fwd(const float* f, const float* cam_norm, float* cam_rv, int N, int C, int K, int h,
                             int w, void* workspace, size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(f && cam_norm && cam_rv && workspace, WSEGK_ERR_NULL, "pcm_fwd: null pointer");
    if (int rc = pcm_check("pcm_fwd", N, C, K, h, w)) return rc;
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, WSEGK_ERR_BAD_ALIGN,
                  "pcm_fwd: workspace must be 1024-byte aligned");
    PcmPlan pl = make_pcm_plan(workspace, N, C, K, h, w);
    WSEGK_REQUIRE(workspace_bytes >= pl.bytes, WSEGK_ERR_WORKSPACE_TOO_SMALL, "pcm_fwd: workspace %zu < %zu",
                  workspace_bytes, pl.bytes);
    {
        const size_t smem = static_cast<size_t>(K) * 129 * sizeof(float);
        WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_prep_f_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        pcm_prep_f_kernel<<<dim3(pl.T, N), 256, smem, stream>>>(f, pl.fhat, pl.inv_norm, K, pl.HW, pl.T, pl.HWp);
        pcm_prep_cam_kernel<<<dim3(pl.T, N), 128, 0, stream>>>(cam_norm, pl.camb, C, pl.HW, pl.T);
    }
    {
        const size_t smem = fwd_smem_bytes(K);
        WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        prof_mark(0, stream);
        pcm_fwd_kernel<<<dim3(pl.T, pl.nsplit, N), kFwdThreads, smem, stream>>>(pl.fhat, pl.camb, pl.part, K,
                                                                                pl.T, pl.HWp, pl.nsplit);
        prof_mark(1, stream);
    }
    pcm_fwd_finalize_kernel<<<dim3(ceil_div(pl.HWp, 128), N), 128, 0, stream>>>(pl.part, cam_rv, pl.dsum, C,
                                                                                 pl.HW, pl.HWp, pl.nsplit);
    count_launches(4);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_bwd(const float* f, const float* cam_norm, const float* cam_rv, const float* grad_cam_rv,
                             float* grad_f, int N, int C, int K, int h, int w, void* workspace,
                             size_t workspace_bytes, wsegk_stream_t stream) {
    if (int rc = check_device()) return rc;
    WSEGK_REQUIRE(f && cam_norm && cam_rv && grad_cam_rv && grad_f && workspace, WSEGK_ERR_NULL,
                  "pcm_bwd: null pointer");
    if (int rc = pcm_check("pcm_bwd", N, C, K, h, w)) return rc;
    WSEGK_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 1023) == 0, WSEGK_ERR_BAD_ALIGN,
                  "pcm_bwd: workspace must be 1024-byte aligned");
    PcmPlan pl = make_pcm_plan(workspace, N, C, K, h, w);
    WSEGK_REQUIRE(workspace_bytes >= pl.bytes, WSEGK_ERR_WORKSPACE_TOO_SMALL, "pcm_bwd: workspace %zu < %zu",
                  workspace_bytes, pl.bytes);
    pcm_bwd_prep_kernel<<<dim3(pl.T, N), 128, 0, stream>>>(cam_norm, cam_rv, grad_cam_rv, pl.dsum, pl.uaug,
                                                           pl.vaug, C, pl.HW, pl.T, pl.HWp);
    {
        const size_t smem = bwd_smem_bytes(K);
        WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        prof_mark(2, stream);
        pcm_bwd_kernel<<<dim3(pl.T, pl.nsplit, N), kFwdThreads, smem, stream>>>(pl.fhat, pl.uaug, pl.vaug,
                                                                                pl.dfhat, K, pl.T, pl.HWp,
                                                                                pl.nsplit);
        prof_mark(3, stream);
    }
    {
        const size_t smem = static_cast<size_t>(K) * 129 * sizeof(float);
        WSEGK_CUDA_OK(cudaFuncSetAttribute(pcm_bwd_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           static_cast<int>(smem)));
        pcm_bwd_post_kernel<<<dim3(pl.T, N), 256, smem, stream>>>(pl.dfhat, f, pl.inv_norm, grad_f, K, pl.HW,
                                                                  pl.HWp, pl.nsplit);
    }
    count_launches(3);
    WSEGK_CUDA_OK(cudaGetLastError());
    return WSEGK_OK;
}
