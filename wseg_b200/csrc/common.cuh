// Shared host-side helpers for the wsegk C ABI: status codes, thread-local error text, device gate,
// workspace carving.  No torch types anywhere below the ABI.
#pragma once
#include <cstdarg>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#include "../../include/wsegk.h"

namespace wsegk {

constexpr int kTile = 128;        // pixels per A-side (j) PCM tile = rows one CTA owns (UMMA M = 256 over the CTA pair)
#ifndef WSEGK_CAM_ROWS
#define WSEGK_CAM_ROWS 32
#endif
// PCM forward: cam operand rows per CTA (rank 0: cam hi + the ones row, rank 1: cam lo).  -DWSEGK_CAM_ROWS=24 (the smallest
// multiple of 8 that holds kMaxClasses + 1 rows) makes the P V product 256 x 48 x 16 instead of 256 x 64 x 16: 6 % less
// tensor-pipe work per stage, parity-green, and measured no faster (forward 47-49 us either way on 52 pairs: the kernel is
// not quite pipe-bound), so the profiled configuration stays.
constexpr int kCamRows = WSEGK_CAM_ROWS;
constexpr int kONum = 2 * kCamRows;   // forward accumulator columns = rows of a partial slot
static_assert(kCamRows % 8 == 0 && kCamRows >= 24 && kCamRows <= 32, "cam operand rows");
constexpr int kAugK = 48;         // PCM backward rank-(2C+2) factor, padded to a multiple of 16
constexpr int kMaxClasses = 23;   // 2C + 2 <= kAugK
constexpr int kMaxPcmK = 192;
constexpr int kMaskWords = 4;     // sign-mask words per (stage, row): up to 128 columns
constexpr int kBoundLo2 = 8;      // slot of PcmPlan::bound holding the largest squared fp16 rounding residual of a pixel
// PCM call flags (include/wsegk.h: WSEGK_PCM_*)
constexpr int kPcmPrecise = 1, kPcmDeferFixup = 2;

void set_error(const char* fmt, ...);
void count_launches(int n);   // diagnostics: wsegk_debug_launch_count()
int check_device();   // WSEGK_OK if the current device is sm_100, else WSEGK_ERR_DEVICE (message set)

#define WSEGK_CUDA_OK(expr)                                                                    \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            ::wsegk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                               __LINE__);                                                      \
            return WSEGK_ERR_CUDA;                                                             \
        }                                                                                      \
    } while (0)

#define WSEGK_REQUIRE(cond, code, ...)      \
    do {                                    \
        if (!(cond)) {                      \
            ::wsegk::set_error(__VA_ARGS__); \
            return (code);                  \
        }                                   \
    } while (0)

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }
inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline int ceil_div_dev(int a, int b) { return (a + b - 1) / b; }
size_t fullres_workspace_bytes(int N, int C);
size_t lazy_bwd_workspace_bytes(int N, int C, int H, int W);
// gx[plane] = transpose of the bilinear enlargement applied to gy[plane]; planes whose label is 0 (and, if c0_stride > 0,
// planes with plane % c0_stride == 0) are not read: their gx is zero
int launch_resize_bwd_labelled(const float* gy, const float* label, float* gx, int planes, int H, int W, int oh, int ow,
                               int c0_stride, cudaStream_t stream);
size_t proto_workspace_bytes(int N, int C, int h, int w);
size_t contrast_workspace_bytes(int N, int C, int D, int h, int w);
size_t select_workspace_bytes(int rows);

// Bump allocator over the caller's workspace; every section is 1 KiB aligned.
struct Carver {
    uint8_t* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<uint8_t*>(p)) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 1024);
        T* p = reinterpret_cast<T*>(reinterpret_cast<uintptr_t>(base) + off);   // base may be null: offsets only
        off += count * sizeof(T);
        return p;
    }
    size_t total() const { return align_up(off, 1024); }
};

// ---- PCM workspace (shared by forward and backward) -------------------------------------------
// A CTA pair owns 256 consecutive "j" pixels (two 128-row A tiles) and sweeps the "i" pixels in stages of
// 2*wb pixels (one wb-row B tile per CTA).  All operand tiles are stored pre-tiled in the UMMA canonical
// no-swizzle K-major layout so each is ONE contiguous bulk copy:
//   chunk (row r, 8-element K group g) of an R-row tile lives at byte (g*R + r)*16.
struct PcmPlan {
    int N, C, K, HW;
    int PJ;          // CTA pairs along j:  ceil(HW / 256)
    int HWpj;        // PJ * 256
    int wb;          // B-tile rows per CTA (64, 56 or 48)
    int NS;          // stages along i: ceil(HW / (2*wb))
    int HWpi;        // NS * 2 * wb
    int G, Gb;       // CTA pairs of the persistent forward / backward main kernels (stream-K over all N*PJ*NS stages)
    int nsplit;      // partial-result slots per unit = the most CTA pairs that share one (image, j pair) sweep
    int precise;     // 1: exact sign mask (near-zero entries of S re-evaluated in fp32) + split backward factors; 0: single pass
    size_t a_tile_bytes, b_tile_bytes;
    // sections
    uint16_t* fa;                          // [N][2*PJ] A tiles (128 rows) of the normalised features, fp16
    uint16_t* fb;                          // [N][2*NS] B tiles (wb rows) of the same
    float* fh32;                           // [N][HWpj][K] normalised features, fp32, pixel-major (precise mode: sign fix-up)
    uint32_t* nzmask;                      // [N][NS][HWpj][4] bit = |S| below the fp16-rounding threshold (precise mode)
    uint16_t* camb;                        // [N][NS][2 ranks][2wb/8][32 rows][8] fp16 cam operand blocks
    float* inv_norm;                       // [N][HWpj]   1/(|f|+1e-5)
    float* dsum;                           // [N][HWpj]   column sums of relu(S) + 1e-5
    float* part;                           // [N][nsplit][HWpj][64] forward partial numerators
    uint32_t* mask;                        // [N][NS][HWpj][4] sign bits of S (row j, stage columns)
    uint16_t* ua;                          // [N][2*PJ] A tiles [128 x Ku] bf16 backward factor U
    uint16_t* vb;                          // [N][2*NS] B tiles [wb x Ku] bf16 backward factor V
    float* dfhat;                          // [N][nsplit][K][HWpj] fp32 partial gradient w.r.t. normalised features
    float* bound;                          // [0] upper bound of |T|, [kBoundLo2] max_p |fh_p - fp16(fh_p)|^2 (float bits, atomicMax)
    float* peak;                           // [N][C] per-plane maxima of relu(cam) + 1e-5 (pcm_fwd_logits: fused K1a)
    float* out;                            // forward: the caller's cam_rv [N][C][HW] (not part of the workspace)
    int fwd_shared_units;                  // forward schedule: some unit is swept by more than one CTA pair (finalize kernel needed)
    int Ku;                                // 48 (fast) or 144 (precise: [hi|hi|lo] x [hi|lo|hi])
    size_t bytes;
};
// ---- schedule shared by the forward and backward main kernels: data-parallel rows + a stream-K tail ------------------
// Unit u = n*PJ + jp (one image's pair of 128-row j tiles) owns NS stages.  G CTA pairs run persistently:
//   * rows:  for k < Rf = U / G, pair g runs the whole unit k*G + g.  In step k all pairs sit on consecutive units, i.e. on
//            the same few images, so the i-side tiles they stream are shared through L2;
//   * tail:  the remaining r = U - Rf*G units have Tt = r*NS stages, cut into Gt = min(G, Tt) equal contiguous ranges, one
//            per pair g < Gt (stream-K): every pair ends up with the same number of stages (+-1) whatever N, PJ and NS are.
// A range crosses unit boundaries; each (pair, unit) piece is a "segment" with its own A-side tile and its own
// partial-result slot: slot 0 for a row unit, pair - (pair owning the unit's first tail stage) for a tail unit.
struct PcmSched {
    int U, NS, G, Rf, r, Gt;
    long long Tt;
};
__host__ __device__ __forceinline__ PcmSched pcm_sched(int U, int NS, int G) {
    PcmSched s;
    s.U = U; s.NS = NS; s.G = G;
    s.Rf = U / G;
    s.r = U - s.Rf * G;
    s.Tt = static_cast<long long>(s.r) * NS;
    s.Gt = s.Tt < G ? static_cast<int>(s.Tt) : G;
    return s;
}
__host__ __device__ __forceinline__ long long pcm_tail_lo(const PcmSched& s, int g) {       // g <= Gt
    return s.Gt > 0 ? s.Tt * g / s.Gt : 0;
}
__host__ __device__ __forceinline__ int pcm_tail_owner(const PcmSched& s, long long t) {    // Gt <= Tt
    const int g = static_cast<int>(t * s.Gt / s.Tt);
    return pcm_tail_lo(s, g + 1) > t ? g : g + 1;
}
// segments (= partial slots in use) of unit u
__host__ __device__ __forceinline__ int pcm_unit_segments(const PcmSched& s, int u) {
    const int w = u - s.Rf * s.G;
    if (w < 0) return 1;
    const long long t0 = static_cast<long long>(w) * s.NS;
    return pcm_tail_owner(s, t0 + s.NS - 1) - pcm_tail_owner(s, t0) + 1;
}

// Walks the stages of a CTA pair's schedule (common.cuh) without per-stage divisions: the MMA-issuing warp has only
// ~500 spare cycles per stage.
struct StageCursor {
    int n, jp, st, k;          // image, j pair, stage inside the unit, segment ordinal inside this CTA pair's schedule
    int i, cnt;                // running stage index, stages in the schedule
    int u, rows_left, tail_u, tail_st, tail_i, G, NS, PJ;
    __host__ __device__ __forceinline__ void set_unit(int unit) {
        u = unit;
        n = unit / PJ;
        jp = unit - n * PJ;
    }
    __host__ __device__ __forceinline__ StageCursor(const PcmSched& sc, int g, int PJ_) : k(0), i(0), G(sc.G), NS(sc.NS), PJ(PJ_) {
        const long long lo = g < sc.Gt ? pcm_tail_lo(sc, g) : 0, hi = g < sc.Gt ? pcm_tail_lo(sc, g + 1) : 0;
        rows_left = sc.Rf;
        tail_i = sc.Rf * sc.NS;                                      // index of the first tail stage
        cnt = tail_i + static_cast<int>(hi - lo);
        const int w = static_cast<int>(lo / sc.NS);
        tail_u = sc.Rf * sc.G + w;
        tail_st = static_cast<int>(lo - static_cast<long long>(w) * sc.NS);
        if (rows_left > 0) { set_unit(g); st = 0; }
        else { set_unit(tail_u < sc.U ? tail_u : 0); st = tail_st; }
    }
    __host__ __device__ __forceinline__ bool first() const { return i == 0 || st == 0 || i == tail_i; }   // first stage of its segment
    __host__ __device__ __forceinline__ bool last() const { return i + 1 == cnt || st == NS - 1; }         // last stage of its segment
    __host__ __device__ __forceinline__ bool in_tail() const { return i >= tail_i; }
    __host__ __device__ __forceinline__ void next() {
        ++i;
        if (i >= cnt) return;
        if (++st == NS || i == tail_i) {
            ++k;
            if (i == tail_i) { set_unit(tail_u); st = tail_st; }     // rows done: jump to this pair's tail range
            else if (i < tail_i) { st = 0; set_unit(u + G); }        // next row
            else { st = 0; set_unit(u + 1); }                        // next tail unit
        }
    }
};

// peak[plane] = max relu(cam[plane]) + 1e-5, one CTA per (n,c) plane (cam_norm.cu)
void launch_cam_plane_peak(const float* cam, float* peak, int planes, int HW, cudaStream_t stream);

// wb_force / pairs_force > 0: that B-tile height / exactly that many forward CTA pairs (the multi-problem forward, whose
// problems share one kernel instantiation and split the grid between them) instead of the per-shape choices
PcmPlan make_pcm_plan(void* ws, int N, int C, int K, int h, int w, int precise, int wb_force = 0, int pairs_force = 0);
int pcm_mode();     // current process-wide mode (wsegk_pcm_set_mode): 1 precise (default), 0 fast

}  // namespace wsegk
