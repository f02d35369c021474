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

constexpr int kTile = 128;        // pixels per PCM tile (UMMA M and N)
constexpr int kCamRows = 48;      // PCM forward B-operand rows: 24 "hi" rows (21 classes + ones row) + 24 "lo"
constexpr int kCamHalf = 24;
constexpr int kAugRows = 48;      // PCM backward rank-44 factor, padded
constexpr int kMaxClasses = 23;   // C + ones row must fit kCamHalf
constexpr int kMaxPcmK = 192;

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
size_t proto_workspace_bytes(int N, int C, int h, int w);
size_t contrast_workspace_bytes(int N, int C, int D, int h, int w);

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
struct PcmPlan {
    int N, C, K, HW, T, HWp, nsplit;
    size_t tile_bytes;       // one 128-pixel x K bf16 operand tile
    // sections
    uint16_t* fhat;          // [N][T] tiles, canonical no-swizzle K-major layout, bf16
    uint16_t* camb;          // [N][T] tiles of [48 x 128] bf16 (cam hi / ones / cam lo), K-major
    float* inv_norm;         // [N][HWp]   1/(|f|+1e-5)
    float* dsum;             // [N][HWp]   column sums of relu(S) + 1e-5
    float* part;             // [N][nsplit][HWp][48] forward partial numerators
    uint16_t* uaug;          // [N][T] tiles of [128 x 48] bf16   backward factor U (K-major)
    uint16_t* vaug;          // [N][T] tiles of [128 x 48] bf16   backward factor V (K-major)
    float* dfhat;            // [N][HWp][K] fp32 gradient w.r.t. normalised features (pixel-major)
    size_t bytes;
};
PcmPlan make_pcm_plan(void* ws, int N, int C, int K, int h, int w);
int pcm_split(int N, int T);

}  // namespace wsegk
