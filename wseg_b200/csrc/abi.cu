// wsegk C ABI: version, error text, device gate, workspace sizing.
#include <algorithm>
#include <atomic>
#include <cstdlib>
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

// default mode of the entry points that do not take flags: per HOST THREAD, so that concurrent callers (DataParallel's
// parallel_apply, the thread pool of contrast_infer.py) cannot change each other's mode
static thread_local int g_pcm_mode = 1;
int pcm_mode() { return g_pcm_mode; }

// B-tile rows per CTA: the candidate that pads the i sweep least (ties: the larger tile, fewer stages).
// WSEGK_WB48_SLACK_PCT >= 0 (build flag or environment; A/B aid) prefers 48-row tiles whenever they pad at most that many per
// cent more than the best: only with 96-column stages do three T / E buffers fit beside the backward's accumulator in TMEM
// (pcm.cu: BwdTmem).  Measured round 3 at N = 8, 56x56: the backward does NOT get faster with the third buffer (61.5 us both
// ways — its TMEM warps, not the buffer ring, bound it) and the forward loses 4 us to the 18 % more stages: off by default.
#ifndef WSEGK_WB48_SLACK_PCT
#define WSEGK_WB48_SLACK_PCT (-1)
#endif
static int wb48_slack_pct() {
    static const int v = [] { const char* e = getenv("WSEGK_WB48_SLACK_PCT"); return e ? atoi(e) : WSEGK_WB48_SLACK_PCT; }();
    return v;
}
static int pick_wb(int HW) {
    const int cands[3] = {64, 56, 48};
    int best = 64;
    long long best_pad = -1, pad48 = 0;
    for (int wb : cands) {
        const long long ns = (HW + 2 * wb - 1) / (2 * wb);
        const long long padded = ns * 2 * wb;
        if (wb == 48) pad48 = padded;
        if (best_pad < 0 || padded < best_pad) { best_pad = padded; best = wb; }
    }
    if (wb48_slack_pct() >= 0 && pad48 * 100 <= best_pad * (100 + wb48_slack_pct())) return 48;
    {
        // the same switch for 64-row tiles (128-column stages: fewer stages, full-width MMAs, but only three S buffers)
        static const int s64 = [] { const char* e = getenv("WSEGK_WB64_SLACK_PCT"); return e ? atoi(e) : -1; }();
        const long long pad64 = ((HW + 127) / 128) * 128;
        if (s64 >= 0 && pad64 * 100 <= best_pad * (100 + s64)) return 64;
    }
    return best;
}

// How many CTA pairs split one j pair's sweep over the i stages: minimise (waves x stages per CTA + fixed cost).
// CTA pairs of the persistent PCM kernels.  A B200 has 74 SM pairs and a PCM CTA takes a whole SM (213 KB of shared memory),
// so 74 pairs is the fastest PCM in isolation — but in the training step the other chains (K1b, seeds, prototypes, NCE) then
// cannot make progress while a PCM kernel runs.  72 pairs leave 4 SMs to them: the PCM kernels get ~3 % slower, the whole
// step 1.7 % faster (A/B on one box, 68 / 70 / 72 within noise of each other, 66 and 74 slower; profiles/README.md).
// A constant, not a device query, so that workspace sizes do not depend on where they are asked for.
// Defaults tuned on the whole training step (A/B on one box, profiles/README.md round 2): a B200 has 74 SM pairs and a PCM
// CTA takes a whole SM, so every pair a PCM kernel does not use is two SMs on which the HBM- and latency-bound kernels of the
// step's other chains run meanwhile.  In isolation 72 pairs are fastest (forward 46 us, backward 54 us at N = 8, 56x56); in
// the step 64 / 52 pairs are (0.334 -> 0.31 ms; anything in 52..64 / 48..56 is within noise) although the two kernels alone
// then take 51 / 64 us.  52 divides the 104 units of the headline shape: the backward then has no stream-K tail and one partial slot.
#ifndef WSEGK_PCM_PAIRS_FWD
#define WSEGK_PCM_PAIRS_FWD 64
#endif
#ifndef WSEGK_PCM_PAIRS_BWD
#define WSEGK_PCM_PAIRS_BWD 52
#endif
// WSEGK_PCM_PAIRS_FWD / _BWD in the environment override the built-in defaults (A/B runs of the whole step)
static int default_pairs(int which) {
    static const int v[2] = {[] { const char* e = getenv("WSEGK_PCM_PAIRS_FWD"); const int x = e ? atoi(e) : 0;
                                  return x >= 1 && x <= 74 ? x : WSEGK_PCM_PAIRS_FWD; }(),
                             [] { const char* e = getenv("WSEGK_PCM_PAIRS_BWD"); const int x = e ? atoi(e) : 0;
                                  return x >= 1 && x <= 74 ? x : WSEGK_PCM_PAIRS_BWD; }()};
    return v[which];
}
static thread_local int g_pcm_pairs[2] = {default_pairs(0), default_pairs(1)};

extern "C" int wsegk_pcm_set_pairs(int fwd_pairs, int bwd_pairs) {
    if (fwd_pairs < 0 || fwd_pairs > 74 || bwd_pairs < 0 || bwd_pairs > 74) {
        set_error("pcm_set_pairs: need 0 (default) .. 74 CTA pairs, got %d / %d", fwd_pairs, bwd_pairs);
        return WSEGK_ERR_BAD_SHAPE;
    }
    g_pcm_pairs[0] = fwd_pairs ? fwd_pairs : default_pairs(0);
    g_pcm_pairs[1] = bwd_pairs ? bwd_pairs : default_pairs(1);
    return WSEGK_OK;
}

// U units of NS stages on at most `slots` CTA pairs.  If U < 3 slots and a pair count in [3/4 slots, slots] divides U, the
// largest such is taken: every unit is then swept by ONE pair — no stream-K tail, no partial slots, and the forward needs no finalize pass
// (104 units of the headline shape: 52 pairs instead of 64; the kernel alone 51 -> 57 us, the finalize kernel's 10 us and
// its launch leave the chain, and 24 more SMs stay free for the step's other kernels).
static int pick_clusters(long long U, long long NS, long long slots) {
    long long g = U * NS / 2;
    if (g > slots) g = slots;
    if (g < 1) g = 1;
#ifndef WSEGK_PCM_NO_DIVISOR
    if (g == slots && U > g && U < 3 * g)      // only while a unit is a large part of a pair's work: with many units per pair the
        for (long long d = g; 4 * d >= 3 * g; --d)   // stream-K tail is a small share and fewer pairs only cost time
            if (U % d == 0) return static_cast<int>(d);
#endif
    return static_cast<int>(g);
}
static int max_unit_segments(int units, int NS, int G) {
    const PcmSched sc = pcm_sched(units, NS, G);
    int best = 1;
    for (int u = sc.Rf * sc.G; u < units; ++u) best = std::max(best, pcm_unit_segments(sc, u));
    return best;
}

PcmPlan make_pcm_plan(void* ws, int N, int C, int K, int h, int w, int precise, int wb_force, int pairs_force) {
    PcmPlan p{};
    p.N = N; p.C = C; p.K = K; p.HW = h * w;
    p.PJ = ceil_div(p.HW, 2 * kTile);
    p.HWpj = p.PJ * 2 * kTile;
    p.wb = wb_force > 0 ? wb_force : pick_wb(p.HW);
    p.NS = ceil_div(p.HW, 2 * p.wb);
    p.HWpi = p.NS * 2 * p.wb;
    // forward and backward main kernels may run on different numbers of CTA pairs (round 2: with the step re-ordered the
    // optimum moved down — what the other chains gain from the SMs a PCM kernel leaves free outweighs what the PCM loses)
    p.G = pick_clusters(static_cast<long long>(N) * p.PJ, p.NS, g_pcm_pairs[0]);
    p.Gb = pick_clusters(static_cast<long long>(N) * p.PJ, p.NS, g_pcm_pairs[1]);
    if (pairs_force > 0) p.G = p.Gb = pairs_force;
    const int seg_f = max_unit_segments(N * p.PJ, p.NS, p.G);
    p.fwd_shared_units = seg_f > 1;
    p.nsplit = std::max(seg_f, max_unit_segments(N * p.PJ, p.NS, p.Gb));
    p.precise = precise ? 1 : 0;
    p.Ku = precise ? 3 * kAugK : kAugK;
    p.a_tile_bytes = static_cast<size_t>(kTile) * K * 2;
    p.b_tile_bytes = static_cast<size_t>(p.wb) * K * 2;
    Carver c(ws);
    const size_t at = static_cast<size_t>(N) * 2 * p.PJ, bt = static_cast<size_t>(N) * 2 * p.NS;
    p.fa = c.take<uint16_t>(at * kTile * K);
    p.fh32 = c.take<float>(static_cast<size_t>(N) * p.HWpj * K);
    p.fb = c.take<uint16_t>(bt * p.wb * K);
    p.nzmask = c.take<uint32_t>(static_cast<size_t>(N) * p.NS * p.HWpj * kMaskWords);
    p.camb = c.take<uint16_t>(static_cast<size_t>(N) * p.NS * 2 * (2 * p.wb) * kCamRows);
    p.inv_norm = c.take<float>(static_cast<size_t>(N) * p.HWpj);
    p.dsum = c.take<float>(static_cast<size_t>(N) * p.HWpj);
    p.part = c.take<float>(static_cast<size_t>(N) * p.nsplit * p.HWpj * kONum);
    p.mask = c.take<uint32_t>(static_cast<size_t>(N) * p.NS * p.HWpj * kMaskWords);
    p.ua = c.take<uint16_t>(at * kTile * 3 * kAugK);
    p.vb = c.take<uint16_t>(bt * p.wb * 3 * kAugK);
    p.dfhat = c.take<float>(static_cast<size_t>(N) * p.nsplit * K * p.HWpj);
    p.bound = c.take<float>(256);
    p.peak = c.take<float>(static_cast<size_t>(N) * C);
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
            return make_pcm_plan(nullptr, N, C, K, h, w, 1).bytes;   // mode-independent (sized for precise)
        case WSEGK_OP_FULLRES_NORM:
            return fullres_workspace_bytes(N, C);
        case WSEGK_OP_LAZY_BWD:
            return lazy_bwd_workspace_bytes(N, C, h, w);
        case WSEGK_OP_SEED_LABELS:
            return align_up(static_cast<size_t>(N) * C * 2 * sizeof(float), 1024);
        case WSEGK_OP_PROTO:
            return proto_workspace_bytes(N, C, h, w);
        case WSEGK_OP_CONTRAST_LOSS:
            return contrast_workspace_bytes(N, C, D, h, w);
        case WSEGK_OP_SELECT:          // N = rows (also what minpool_fwd needs)
            return select_workspace_bytes(N);
        case WSEGK_OP_ER_ECR:          // N images on an h x w grid
            return align_up(static_cast<size_t>(N) * ceil_div(h * w, 128) * sizeof(float), 256) + select_workspace_bytes(N);
        default:
            set_error("unknown op %d", op);
            return 0;
    }
}

extern "C" int wsegk_pcm_set_mode(int mode) {
    if (mode != 0 && mode != 1) {
        set_error("pcm_set_mode: mode must be 0 (fast) or 1 (precise), got %d", mode);
        return WSEGK_ERR_BAD_SHAPE;
    }
    g_pcm_mode = mode;
    return WSEGK_OK;
}

extern "C" int wsegk_pcm_get_mode(void) { return pcm_mode(); }

extern "C" int wsegk_debug_pcm_layout(int N, int C, int K, int h, int w, size_t* out, int n_out) {
    if (!out || N <= 0 || h <= 0 || w <= 0) return 0;
    PcmPlan p = make_pcm_plan(nullptr, N, C, K, h, w, 1);
    auto off = [](const void* q) { return reinterpret_cast<size_t>(q); };   // base is nullptr => offsets
    const size_t v[20] = {off(p.fa), off(p.fh32), off(p.fb), off(p.nzmask), off(p.camb), off(p.inv_norm),
                          off(p.dsum), off(p.part), off(p.mask), off(p.ua), off(p.vb), off(p.dfhat), off(p.bound),
                          p.bytes, static_cast<size_t>(p.nsplit), static_cast<size_t>(p.PJ),
                          static_cast<size_t>(p.HWpj), static_cast<size_t>(p.wb), static_cast<size_t>(p.NS),
                          static_cast<size_t>(p.HWpi)};
    int k = 0;
    for (; k < n_out && k < 20; ++k) out[k] = v[k];
    return k;
}

// Test aid (host only): the stages CTA pair g runs under the schedule of the PCM main kernels for N images on an h x w grid,
// as (unit, stage, segment ordinal, partial slot) quadruples; returns how many, or a negative status.
extern "C" int wsegk_debug_pcm_schedule(int N, int h, int w, int g, int* out, int capacity) {
    if (N <= 0 || h <= 0 || w <= 0 || g < 0) return WSEGK_ERR_BAD_SHAPE;
    const PcmPlan pl = make_pcm_plan(nullptr, N, 1, 32, h, w, 1);
    if (g >= pl.G) return 0;
    const PcmSched sc = pcm_sched(N * pl.PJ, pl.NS, pl.G);
    StageCursor c(sc, g, pl.PJ);
    const int cnt = c.cnt;
    for (int i = 0; i < cnt; ++i, c.next()) {
        if (out != nullptr && i < capacity) {
            const int slot = c.in_tail() ? g - pcm_tail_owner(sc, static_cast<long long>(c.u - sc.Rf * sc.G) * sc.NS) : 0;
            out[4 * i + 0] = c.u;
            out[4 * i + 1] = c.st;
            out[4 * i + 2] = c.k;
            out[4 * i + 3] = slot;
        }
    }
    return cnt;
}
