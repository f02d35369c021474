// wsegk C ABI: version, error text, device gate, workspace sizing.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace wsegk {

static thread_local char g_err[512] = "";
static std::atomic<unsigned long long> g_launches{0};

void count_launches(int n) { g_launches.fetch_add(static_cast<unsigned long long>(n), std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_device() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        set_error("no CUDA device: %s (wsegk has no CPU path)", cudaGetErrorString(e));
        return WSEGK_ERR_DEVICE;
    }
    static std::mutex mu;
    static int cached_major[64];
    static bool have[64];
    int major = 0;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 64 && have[dev]) {
            major = cached_major[dev];
        } else {
            if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) {
                set_error("cudaDeviceGetAttribute failed");
                return WSEGK_ERR_DEVICE;
            }
            if (dev < 64) {
                cached_major[dev] = major;
                have[dev] = true;
            }
        }
    }
    if (major != 10) {
        set_error("device %d is sm_%dx; wsegk kernels are built for sm_100a only (no fallback)", dev, major);
        return WSEGK_ERR_DEVICE;
    }
    return WSEGK_OK;
}

// How many CTAs split one j-tile's loop over i-tiles, chosen so the grid covers the 148 SMs about twice.
int pcm_split(int N, int T) {
    int ctas = std::max(1, N * T);
    int s = (2 * 148 + ctas - 1) / ctas;
    s = std::max(1, std::min(s, std::min(T, 8)));
    return s;
}

PcmPlan make_pcm_plan(void* ws, int N, int C, int K, int h, int w) {
    PcmPlan p{};
    p.N = N; p.C = C; p.K = K; p.HW = h * w;
    p.T = ceil_div(p.HW, kTile);
    p.HWp = p.T * kTile;
    p.nsplit = pcm_split(N, p.T);
    p.tile_bytes = static_cast<size_t>(kTile) * K * 2;
    Carver c(ws);
    const size_t tiles = static_cast<size_t>(N) * p.T;
    p.fhat = c.take<uint16_t>(tiles * kTile * K);
    p.camb = c.take<uint16_t>(tiles * kCamRows * kTile);
    p.inv_norm = c.take<float>(static_cast<size_t>(N) * p.HWp);
    p.dsum = c.take<float>(static_cast<size_t>(N) * p.HWp);
    p.part = c.take<float>(static_cast<size_t>(N) * p.nsplit * p.HWp * kCamRows);
    p.uaug = c.take<uint16_t>(tiles * kTile * kAugRows);
    p.vaug = c.take<uint16_t>(tiles * kTile * kAugRows);
    p.dfhat = c.take<float>(static_cast<size_t>(N) * p.nsplit * p.HWp * K);
    p.bytes = c.total();
    return p;
}

}  // namespace wsegk

using namespace wsegk;

extern "C" int wsegk_version(void) { return 100; }   // 0.1.0

extern "C" const char* wsegk_last_error(void) { return g_err; }

extern "C" unsigned long long wsegk_debug_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" size_t wsegk_workspace_bytes(int op, int N, int C, int K, int D, int h, int w) {
    if (N <= 0 || h <= 0 || w <= 0) return 0;
    switch (op) {
        case WSEGK_OP_CAM_NORM:
            return align_up(static_cast<size_t>(N) * C * sizeof(float), 1024);
        case WSEGK_OP_PCM:
            return make_pcm_plan(nullptr, N, C, K, h, w).bytes;
        case WSEGK_OP_FULLRES_NORM:
            return fullres_workspace_bytes(N, C);
        case WSEGK_OP_SEED_LABELS:
            return align_up(static_cast<size_t>(N) * C * 2 * sizeof(float), 1024);
        case WSEGK_OP_PROTO:
            return proto_workspace_bytes(N, C, h, w);
        case WSEGK_OP_CONTRAST_LOSS:
            return contrast_workspace_bytes(N, C, D, h, w);
        default:
            set_error("unknown op %d", op);
            return 0;
    }
}

extern "C" int wsegk_debug_pcm_layout(int N, int C, int K, int h, int w, size_t* out, int n_out) {
    if (!out || N <= 0 || h <= 0 || w <= 0) return 0;
    PcmPlan p = make_pcm_plan(nullptr, N, C, K, h, w);
    auto off = [](const void* q) { return reinterpret_cast<size_t>(q); };   // base is nullptr => offsets
    const size_t v[12] = {off(p.fhat), off(p.camb), off(p.inv_norm), off(p.dsum), off(p.part), off(p.uaug),
                          off(p.vaug), off(p.dfhat), p.bytes, static_cast<size_t>(p.nsplit),
                          static_cast<size_t>(p.T), static_cast<size_t>(p.HWp)};
    int k = 0;
    for (; k < n_out && k < 12; ++k) out[k] = v[k];
    return k;
}
