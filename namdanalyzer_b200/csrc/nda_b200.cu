// nda_b200.cu -- host side of libnda_b200.so: context, buffer cache, launch logic and the C ABI
// declared in include/nda_b200.h.  Compiled for sm_100a only; there is no CPU path.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <stdlib.h>
#include <fcntl.h>
#include <unistd.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <string>
#include <vector>
#include <algorithm>

#include <nvtx3/nvToolsExt.h>
#include <omp.h>
#include <nccl.h>
#include <type_traits>
#include <dlfcn.h>
#include <condition_variable>
#include <functional>
#include <map>

#include "../../include/nda_b200.h"
#include "nda_kernels.cuh"
#include "nda_tc.cuh"
#include "nda_surface.cuh"

// ------------------------------------------------------------------------------------------
// errors, launch accounting, options
// ------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_within_pure{0};
static std::atomic<int> g_last_filter{-1};          // filter of the last pair-kernel launch: 0..5 FP32 filters, 100 tensor core
static std::atomic<int> g_last_collective{0};       // how the last multi-device call merged its partials: 0 none, 1 NCCL, 2 host sum

// Every tuning / debugging switch of the library.  The environment is read ONCE, when the library is
// first used; afterwards only nda_set_option() changes a value (tests and tools use it), so an entry
// point never calls getenv().  Names are the environment variables' names.
struct Config {
    int    filterTc = 1;            // NDA_FILTER=tc (default) | rint | fold | fold2 | fold2a0..3
    int    filterKind = kFilterFold2 + 1;
    double tcRangeFactor = 1.5;     // NDA_TC_RANGE_FACTOR: longest range the tensor filter takes, thr <= factor * P
    double tcMaxFraction = 0.0;     // NDA_TC_MAX_FRACTION: cap on the expected in-range fraction (0 = none)
    int    tcAllowPadding = 0;      // NDA_TC_ALLOW_PADDING: tensor filter even for mostly-padding selections
    int    noCellCache = 0;         // NDA_NO_CELL_CACHE
    int    strictSlow = 0;          // NDA_STRICT_SLOW: IEEE divisions in every strict re-evaluation
    int    tcNoSure = 0;            // NDA_TC_NO_SURE: within without the SURE shortcut
    int    tcTrace = 0;             // NDA_TC_TRACE: clock64 trace buffer (trace builds only, tools/)
    int    ctasPerSM = 0;           // NDA_CTAS_PER_SM: cap on the persistent FP32 kernels' CTAs per SM
    int    debug = 0;               // NDA_DEBUG: 1 filter choice, 2 + host phase clock
    int    hostThreads = 0;         // NDA_HOST_THREADS: host threads gathering pageable rows (0 = automatic)
    int    uploadChunks = 0;        // NDA_UPLOAD_CHUNKS / NDA_WITHIN_CHUNKS: frame chunks per host call (0 = automatic)
    int    withinChunks = 0;
    int    noTaper = 0;             // NDA_NO_TAPER, NDA_TAPER_LEVELS
    int    taperLevels = 2;
    int    copyStreams = 0;         // NDA_COPY_STREAMS: 0 automatic, 1, 2
    int    d2hCopyEngine = 0;       // NDA_D2H_COPY_ENGINE: bitmask back by cudaMemcpyAsync instead of SM stores
    int    collective = 0;          // NDA_COLLECTIVE: 0 automatic (NCCL when it loads), 1 nccl (fail otherwise), 2 host sum
    int    subpassFrames = 0;       // NDA_SUBPASS_FRAMES: device API in frame sub-passes of this size (0 = one pass)
    int    compatCuda = 0;          // NDA_COMPAT_CUDA: getWithin also flags every refSel atom (legacy CUDA build, SURVEY B-1)
    int    nvtx = 1;                // NDA_NVTX=0 switches the NVTX ranges off
    int    sideTeam = 0;            // NDA_SIDE_TEAM=1: the mask scatter runs beside the row gather on a quarter of the team (measured: no gain)
    double denseFraction = 0.25;    // NDA_DENSE_FRACTION: RDF frames whose cut-off sphere exceeds this share of the cell run all-exact (0 = never)
    int    headTaper = 0;           // NDA_HEAD_TAPER=1: small head chunks for pageable inputs (measured SLOWER, see Uploader::begin)
};
static Config g_cfg;
static std::once_flag g_cfg_once;
static std::mutex g_cfg_mutex;

static const char *const kOptionNames[] = {
    "NDA_FILTER", "NDA_TC_RANGE_FACTOR", "NDA_TC_MAX_FRACTION", "NDA_TC_ALLOW_PADDING", "NDA_NO_CELL_CACHE",
    "NDA_STRICT_SLOW", "NDA_TC_NO_SURE", "NDA_TC_TRACE", "NDA_CTAS_PER_SM", "NDA_DEBUG", "NDA_HOST_THREADS",
    "NDA_UPLOAD_CHUNKS", "NDA_WITHIN_CHUNKS", "NDA_NO_TAPER", "NDA_TAPER_LEVELS", "NDA_COPY_STREAMS",
    "NDA_D2H_COPY_ENGINE", "NDA_COLLECTIVE", "NDA_SUBPASS_FRAMES", "NDA_COMPAT_CUDA", "NDA_NVTX", "NDA_HEAD_TAPER", "NDA_SIDE_TEAM", "NDA_DENSE_FRACTION"};

// value == nullptr or "": back to the default
static bool apply_option(Config &c, const char *name, const char *v)
{
    const Config d;
    const bool has = v && *v;
    auto flag = [&](int &dst, int dflt) { dst = has ? (strcmp(v, "0") != 0) : dflt; };     // present and not "0"
    auto num = [&](int &dst, int dflt, int lo, int hi) { dst = dflt; if (has) { const int x = atoi(v); if (x >= lo && x <= hi) dst = x; } };
    if (!strcmp(name, "NDA_FILTER")) {
        c.filterTc = !has || !strcmp(v, "tc");
        c.filterKind = d.filterKind;
        if (has && !strcmp(v, "rint")) c.filterKind = kFilterRint;
        else if (has && !strcmp(v, "fold")) c.filterKind = kFilterFold;
        else if (has && !strncmp(v, "fold2a", 6) && v[6] >= '0' && v[6] <= '3') c.filterKind = kFilterFold2 + (v[6] - '0');
    } else if (!strcmp(name, "NDA_TC_RANGE_FACTOR")) { c.tcRangeFactor = d.tcRangeFactor; if (has && atof(v) > 0.0) c.tcRangeFactor = atof(v); }
    else if (!strcmp(name, "NDA_TC_MAX_FRACTION")) { c.tcMaxFraction = 0.0; if (has && atof(v) > 0.0) c.tcMaxFraction = atof(v); }
    else if (!strcmp(name, "NDA_TC_ALLOW_PADDING")) flag(c.tcAllowPadding, 0);
    else if (!strcmp(name, "NDA_NO_CELL_CACHE")) flag(c.noCellCache, 0);
    else if (!strcmp(name, "NDA_STRICT_SLOW")) flag(c.strictSlow, 0);
    else if (!strcmp(name, "NDA_TC_NO_SURE")) flag(c.tcNoSure, 0);
    else if (!strcmp(name, "NDA_TC_TRACE")) num(c.tcTrace, 0, 0, 1);
    else if (!strcmp(name, "NDA_CTAS_PER_SM")) num(c.ctasPerSM, 0, 1, 32);
    else if (!strcmp(name, "NDA_DEBUG")) num(c.debug, 0, 0, 9);
    else if (!strcmp(name, "NDA_HOST_THREADS")) num(c.hostThreads, 0, 1, 256);
    else if (!strcmp(name, "NDA_UPLOAD_CHUNKS")) num(c.uploadChunks, 0, 1, 1 << 20);
    else if (!strcmp(name, "NDA_WITHIN_CHUNKS")) num(c.withinChunks, 0, 1, 1 << 20);
    else if (!strcmp(name, "NDA_NO_TAPER")) flag(c.noTaper, 0);
    else if (!strcmp(name, "NDA_TAPER_LEVELS")) num(c.taperLevels, 2, 1, 4);
    else if (!strcmp(name, "NDA_COPY_STREAMS")) num(c.copyStreams, 0, 0, 2);
    else if (!strcmp(name, "NDA_D2H_COPY_ENGINE")) flag(c.d2hCopyEngine, 0);
    else if (!strcmp(name, "NDA_COLLECTIVE")) {
        c.collective = 0;
        if (has && !strcmp(v, "nccl")) c.collective = 1;
        else if (has && !strcmp(v, "host")) c.collective = 2;
    }
    else if (!strcmp(name, "NDA_SUBPASS_FRAMES")) num(c.subpassFrames, 0, 1, 1 << 30);
    else if (!strcmp(name, "NDA_COMPAT_CUDA")) flag(c.compatCuda, 0);
    else if (!strcmp(name, "NDA_NVTX")) num(c.nvtx, 1, 0, 1);
    else if (!strcmp(name, "NDA_HEAD_TAPER")) num(c.headTaper, 0, 0, 1);
    else if (!strcmp(name, "NDA_SIDE_TEAM")) num(c.sideTeam, 0, 0, 1);
    else if (!strcmp(name, "NDA_DENSE_FRACTION")) { c.denseFraction = d.denseFraction; if (has && atof(v) >= 0.0) c.denseFraction = atof(v); }
    else return false;
    return true;
}

static void config_init()
{
    std::call_once(g_cfg_once, [] {
        std::lock_guard<std::mutex> lk(g_cfg_mutex);
        for (const char *n : kOptionNames)
            if (const char *e = getenv(n)) apply_option(g_cfg, n, e);
    });
}

// a call works on a snapshot, so an option changed by another thread never splits one call
static Config config_get()
{
    config_init();
    std::lock_guard<std::mutex> lk(g_cfg_mutex);
    return g_cfg;
}

// NVTX ranges around the phases of a call (upload / pack / pair kernel / all-reduce): free when no tool listens
struct NvtxRange {
    bool on;
    explicit NvtxRange(const char *name, bool enabled = true) : on(enabled) { if (on) nvtxRangePushA(name); }
    ~NvtxRange() { if (on) nvtxRangePop(); }
};

static int fail(int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t e_ = (call);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(NDA_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__,  \
                        cudaGetErrorString(e_));                                              \
    } while (0)

#define LAUNCHED()                                                                            \
    do {                                                                                      \
        g_launches.fetch_add(1, std::memory_order_relaxed);                                   \
        CK(cudaGetLastError());                                                               \
    } while (0)

// ------------------------------------------------------------------------------------------
// per-device context: one stream, grow-only device buffers (the reference cudaMallocs and
// frees on every call, lib/cuda/src/getWithin.cu:84-133)
// ------------------------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) { cudaError_t e = cudaFree(p); p = nullptr; cap = 0; if (e != cudaSuccess) return e; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&p, want); }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// rows [srcRow, srcRow + count*stride) of a host array land at uploaded rows [dstRow, dstRow + count)
struct RowRun { int dstRow, srcRow, stride, count; };

struct Ctx {
    int dev = -1;
    bool ready = false;
    std::mutex mu;                                // one call at a time per DEVICE (contexts of different devices run concurrently)
    Config cfg;                                   // option snapshot of the call in progress
    int hostThreads = 1;                          // host threads this call may use (gather, scatter)
    int numSMs = 0;
    size_t smemOptin = 0;
    cudaStream_t stream = nullptr, copyStream = nullptr, copyStream2 = nullptr;
    cudaEvent_t evJoin = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t evCopied[2] = {nullptr, nullptr}, evDone[2] = {nullptr, nullptr}, evBits[2] = {nullptr, nullptr};
    void *pinned = nullptr;                       // pinned host staging for small results
    size_t pinnedCap = 0;
    cudaError_t pinned_reserve(size_t bytes)
    {
        if (bytes <= pinnedCap) return cudaSuccess;
        if (pinned) { cudaFreeHost(pinned); pinned = nullptr; pinnedCap = 0; }
        cudaError_t e = cudaMallocHost(&pinned, bytes + bytes / 4 + 4096);
        if (e == cudaSuccess) pinnedCap = bytes + bytes / 4 + 4096;
        return e;
    }
    // two LANES = two compute streams with their own pass buffers: consecutive frame chunks of the
    // host entry points run on alternating lanes so the head of chunk k+1's persistent kernel fills
    // the tail of chunk k's.  ctl: [0..1] work counter (u32 x2), stats u64 x2 at +16.
    struct Lane {
        DevBuf packA, packB, packAw, packBw, fp, maxAbs, ctl, embA, embB, thrDot;
        cudaStream_t stream = nullptr;
        // recorded after the last kernel that touches the lane's scratch; the next user (any stream) waits on it,
        // so two calls on different caller streams cannot race on the scratch buffers
        cudaEvent_t evBusy = nullptr;
    };
    Lane lane[2];
    cudaEvent_t evSetup = nullptr, evLaneJoin = nullptr, evSetupCopy = nullptr;
    void *setupStage = nullptr;                   // pinned staging for the small per-call arrays (cell, selections)
    size_t setupCap = 0;
    DevBuf raw[2][2];                             // [array][double buffer] raw coordinate chunks
    DevBuf cell, idx1, idx2, gid, counts, bits, sum;
    std::vector<float> hostCell;                  // cell edges read back for the filter choice (device API)
    const float *hostCellPtr = nullptr;           // ... of this device array, frames [hostCellFb, hostCellFe)
    int hostCellFb = 0, hostCellFe = 0, hostCellUses = 0;
    void *stage[2] = {nullptr, nullptr};          // pinned staging for pageable inputs
    size_t stageCap[2] = {0, 0};
    cudaError_t stage_reserve(int b, size_t bytes)
    {
        if (bytes <= stageCap[b]) return cudaSuccess;
        if (stage[b]) { cudaFreeHost(stage[b]); stage[b] = nullptr; stageCap[b] = 0; }
        cudaError_t e = cudaMallocHost(&stage[b], bytes + bytes / 8 + 4096);
        if (e == cudaSuccess) stageCap[b] = bytes + bytes / 8 + 4096;
        return e;
    }
    double stats[4] = {0, 0, 0, 0};
    // row plan of the last within call, reused while the caller passes the same selections again
    // (frame-by-frame and chunk-by-chunk callers do: selection/selParser.py:117-126)
    struct WithinPlan {
        std::vector<int> refSel, outSel, remap, ref2, out2;
        std::vector<RowRun> runs;
        int nAtoms = -1, nRows = 0;
    } plan;
    bool statsPending = false;                    // counters of the last call still sit in lane[].ctl (read on demand)
    // CUDA-event brackets around the pair kernel of every pass of the last entry-point call
    std::vector<cudaEvent_t> kev;
    int kevUsed = 0;
    cudaError_t kev_record(cudaStream_t st)
    {
        if (kevUsed >= 256) return cudaSuccess;
        if ((int)kev.size() <= kevUsed) {
            cudaEvent_t e;
            cudaError_t r = cudaEventCreate(&e);
            if (r != cudaSuccess) return r;
            kev.push_back(e);
        }
        return cudaEventRecord(kev[kevUsed++], st);
    }
};

static Ctx g_ctx[16];
static std::mutex g_ctx_init_mutex;
// devices of the calling thread: [0] serves the device-pointer API; with more than one the host entry
// points shard their frames over all of them (nda_set_devices)
static thread_local std::vector<int> g_devs;
static thread_local int g_dev = -1;
static int default_host_threads();

static int ctx_for(int dev, Ctx **out)
{
    if (dev < 0 || dev >= 16) return fail(NDA_ERR_ARG, "device index %d not supported", dev);
    Ctx &c = g_ctx[dev];
    CK(cudaSetDevice(dev));
    std::lock_guard<std::mutex> lk(g_ctx_init_mutex);
    if (!c.ready) {
        config_init();
        cudaDeviceProp prop;
        CK(cudaGetDeviceProperties(&prop, dev));
        if (prop.major != 10)
            return fail(NDA_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only",
                        dev, prop.major, prop.minor);
        c.dev = dev;
        c.numSMs = prop.multiProcessorCount;
        c.smemOptin = prop.sharedMemPerBlockOptin;
        CK(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        c.lane[0].stream = c.stream;
        CK(cudaStreamCreateWithFlags(&c.lane[1].stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c.evSetup, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.evLaneJoin, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&c.evSetupCopy, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&c.copyStream, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&c.copyStream2, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&c.evJoin, cudaEventDisableTiming));
        CK(cudaEventCreate(&c.ev0));
        CK(cudaEventCreate(&c.ev1));
        for (int i = 0; i < 2; ++i) {
            CK(cudaEventCreateWithFlags(&c.evCopied[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c.evDone[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c.evBits[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&c.lane[i].evBusy, cudaEventDisableTiming));
            CK(c.lane[i].ctl.reserve(256));
            CK(cudaMemset(c.lane[i].ctl.p, 0, 256));
        }
        c.ready = true;
    }
    *out = &c;
    return NDA_OK;
}

// context of the calling thread's (first) device
static int ctx_get(Ctx **out)
{
    if (g_dev < 0) {
        int cur = 0;
        cudaError_t e = cudaGetDevice(&cur);
        if (e != cudaSuccess) return fail(NDA_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
        g_dev = cur;
    }
    return ctx_for(g_dev, out);
}

// gathering pageable rows is memory-bound host work: 16 threads reach ~50 GB/s on the 16-core GPU hosts
// measured (8: 33 GB/s, more threads than cores: slower).  One process per GPU: the ranks of a node share
// its cores, and OpenMP teams that outnumber the cores spin against each other (2 ranks x 16 threads on 16
// cores: 10x slower end to end, measured).  Read once.
static int default_host_threads()
{
    static const int cached = [] {
        const unsigned hw = std::thread::hardware_concurrency();
        int t = (int)std::max(1u, std::min(16u, hw));
        for (const char *name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"}) {
            const char *e = getenv(name);
            const int ranks = e ? atoi(e) : 0;
            if (ranks > 1) { t = std::max(1, std::min(16, (int)hw / ranks)); break; }   // the ranks share the cores, not the cap of 16
        }
        return t;
    }();
    return cached;
}

// every entry point: lock the device's context, snapshot the options; `share` = how many devices of this
// call divide the host's threads among themselves
static inline void begin_call(Ctx &c, int share = 1)
{
    c.kevUsed = 0;
    c.cfg = config_get();
    const int t = c.cfg.hostThreads > 0 ? c.cfg.hostThreads : default_host_threads();
    c.hostThreads = std::max(1, t / std::max(1, share));
}
static inline cudaError_t lane_acquire(Ctx::Lane &L, cudaStream_t st) { return cudaStreamWaitEvent(st, L.evBusy, 0); }
static inline cudaError_t lane_release(Ctx::Lane &L, cudaStream_t st) { return cudaEventRecord(L.evBusy, st); }
static cudaError_t reset_stats(Ctx &c, cudaStream_t st);

// tuning override for experiments: NDA_CTAS_PER_SM=<n> caps the persistent grid at n CTAs per SM
static int tuned_ctas(const Config &cfg, int occ)
{
    if (cfg.ctasPerSM >= 1 && cfg.ctasPerSM < occ) return cfg.ctasPerSM;
    return occ < 1 ? 1 : occ;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ------------------------------------------------------------------------------------------
// device-level drivers (inputs resident, stream-ordered)
// ------------------------------------------------------------------------------------------
static int launch_pack(Ctx &c, cudaStream_t st, const float *src, int nFtot, const int *gather, int n,
                       int nPad, int f0, int nFc, const int *gid, float4 *dst, unsigned *maxAbs,
                       const float *cell, float4 *dstW, int pairLayout = 0, __half *embA = nullptr)
{
    const long long tiles = (long long)(nPad / 32) * ((nFc + 31) / 32);
    if (tiles > 0x7fffffffLL) return fail(NDA_ERR_ARG, "internal: pack tile count overflows");
    const unsigned grid = (unsigned)std::max<long long>(1, std::min<long long>(tiles, (long long)c.numSMs * 5));
    NvtxRange nv("nda:pack_kernel", c.cfg.nvtx);
    pack_kernel<<<grid, 256, 0, st>>>(src, nFtot, gather, n, nPad, f0, nFc, gid, dst, maxAbs, cell, dstW, pairLayout, embA);
    LAUNCHED();
    return NDA_OK;
}

// NDA_FILTER=tc (default): the tensor-core filter (nda_tc.cuh) whenever every frame of the call is
// eligible, the packed-FP32 fold2 filter otherwise.  NDA_FILTER=rint|fold|fold2|fold2a0..3 forces an
// FP32 filter.
// Host-side mirror of prep_tc_kernel with a 10 % margin: true when no frame of [fb, fe) would be
// sent to the all-exact path by the tensor-core filter.  T = in-range limit on r2.
// maxFraction: optional cap on the expected in-range fraction (cut-off sphere over cell volume).  Measured on
// a B200 (tools/density_sweep.py, config-3 geometry) the tensor filter is ahead at every density it is
// eligible for -- 3.6x at 0.02 % of the pairs in range, 2.1x at 2.2 %, 2.0x at 5.4 % -- so the default is no
// cap; NDA_TC_MAX_FRACTION=<f> sets one.
// Longest range the tensor filter takes, as thr <= factor * P (P = sum (c_k / 2 pi)^2; for a cubic cell
// factor 1 is a range of 0.28 c, 1.5 is 0.34 c, 2 is 0.39 c, 3 is 0.48 c).  The chord bound weakens with
// the range (S saturates at 4 P) and the candidate path takes over: measured on small water boxes
// (tools/range_sweep.py, 14.9 A) the tensor filter is 1.44x / 1.2x / 1.24x ahead of the FP32 filter at
// range / c = 0.22 / 0.28 / 0.34 and behind (0.75x, 0.84x) at 0.40 and 0.48.  Results are identical at
// every range; NDA_TC_RANGE_FACTOR (Config::tcRangeFactor) overrides.
// nStream / nResident: the tensor kernels pad the streamed selection to 128 and the resident one to 256
// atoms; with more than ~4x padding the packed-FP32 kernel (5.4x slower per pair) wins.
static bool tc_frames_eligible(const Config &cfg, const float *cell, int fb, int fe, double T, double maxFraction,
                               int nStream, int nResident)
{
    const double k = 0.15915494309189535;
    {
        const double padded = (double)round_up(nStream, 256) * (double)round_up(nResident, 256);
        if (padded > 4.0 * (double)nStream * (double)nResident && !cfg.tcAllowPadding) return false;
    }
    if (cfg.tcMaxFraction > 0.0) maxFraction = cfg.tcMaxFraction;
    for (int f = fb; f < fe; ++f) {
        const double cx = fabs((double)cell[3 * (size_t)f]), cy = fabs((double)cell[3 * (size_t)f + 1]), cz = fabs((double)cell[3 * (size_t)f + 2]);
        if (!(cx > 0.0 && cy > 0.0 && cz > 0.0) || !(cx < 1.0e30 && cy < 1.0e30 && cz < 1.0e30)) return false;
        const double P = (cx * cx + cy * cy + cz * cz) * k * k;
        const double Eb = P * (4.8828125e-04 * 1.01 + 3.814697265625e-06) + 1.0e-6;
        const double thr = T * 1.001 + 1.0e-6;
        if (!(2.5 * Eb * 1.1 <= thr && thr * 1.1 <= cfg.tcRangeFactor * P * 1.2 && P < 2.5e4)) return false;
        const double sphere = 4.1887902047863905 * T * sqrt(T);
        if (sphere > maxFraction * cx * cy * cz) return false;
    }
    return T > 0.0;
}

// cell edges of frames [fb, fe) on the host: the caller's host array when there is one, else a
// read-back of the device array (device API; 12 bytes per frame).  The read-back needs a stream
// synchronisation, so it is cached per (device pointer, frame range) and refreshed every 64th call:
// the host copy only decides WHICH filter runs -- prep_frames / prep_tc_kernel re-check every frame on
// the device and send a frame that does not qualify to the all-exact path -- so a stale copy can cost
// time, never correctness.
static int host_cell_view(Ctx &c, cudaStream_t st, const float *h_cell, const float *d_cell, int fb, int fe, const float **out)
{
    if (h_cell) { *out = h_cell; return NDA_OK; }
    const bool hit = c.hostCellPtr == d_cell && fb >= c.hostCellFb && fe <= c.hostCellFe && ++c.hostCellUses < 64 &&
                     !c.cfg.noCellCache;
    if (!hit) {
        c.hostCell.resize((size_t)3 * fe);
        CK(cudaMemcpyAsync(c.hostCell.data() + 3 * (size_t)fb, d_cell + 3 * (size_t)fb, (size_t)(fe - fb) * 12, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        c.hostCellPtr = d_cell; c.hostCellFb = fb; c.hostCellFe = fe; c.hostCellUses = 0;
    }
    *out = c.hostCell.data();
    return NDA_OK;
}

// frames per pass so that the packed copies stay within a fixed HBM budget
static int frames_per_pass(size_t bytesPerFrame, int nF)
{
    const size_t budget = (size_t)6 << 30;
    size_t f = budget / (bytesPerFrame ? bytesPerFrame : 1);
    if (f < 1) f = 1;
    if (f > (size_t)nF) f = nF;
    return (int)f;
}

// ------------------------------------------------------------------------------------------
// tensor-core filter drivers (nda_tc.cuh): pack -> embed -> prep -> one persistent kernel per pass
// ------------------------------------------------------------------------------------------
static int launch_embed(cudaStream_t st, const float4 *packed, int nPad, int nFc, const float *cell, int f0,
                        const TcFrame *tf, __half *E, int roleA)
{
    dim3 grid(nPad / 128, std::min(nFc, 65535));
    embed_kernel<<<grid, 128, 0, st>>>(packed, nPad, nFc, cell, f0, tf, E, roleA);
    LAUNCHED();
    return NDA_OK;
}

// streamed blocks per work item: >= 16 items per CTA for the static round-robin to balance (the sameSel
// workload is triangular), but an item stays >= minBlocks steps so that the resident chunk is amortised
static void tc_chunking(TcGeom &g, int numSMs, long long framesInPass, int minBlocks, int maxBlocks = 1024)
{
    const long long target = 16LL * numSMs;
    const long long rows = std::max<long long>(1, framesInPass * g.nRChunks);
    const long long chunksWanted = std::max<long long>(1, (target + rows - 1) / rows);
    long long cs = (g.nSBlocks + chunksWanted - 1) / chunksWanted;
    cs = std::min<long long>(maxBlocks, std::max<long long>(minBlocks, std::min<long long>(1024, cs)));
    cs = (cs + 1) & ~1LL;                                   // even, like nSBlocks: every work item is a whole number of
                                                            // two-block stages, so the step parity is 0 at every item start
    g.chunkS = (int)std::min<long long>(cs, std::max(2, g.nSBlocks));
    g.nSChunks = (g.nSBlocks + g.chunkS - 1) / g.chunkS;
    // spread the stages over the ranges evenly: 79 stages in 7 ranges = 2 x 12 + 5 x 11, not 6 x 12 + 7
    // (with static round-robin scheduling a short last range is pure imbalance)
    const int stages = g.nSBlocks / 2, base = stages / g.nSChunks, rem = stages % g.nSChunks;
    g.chunkS = 2 * (base + (rem ? 1 : 0));
    g.nBigChunks = rem ? rem : g.nSChunks;
    if ((g.nSBlocks | g.chunkS) & 1) { g.chunkS = 0; g.nSChunks = 0; }   // cannot happen; launches nothing rather than mis-stepping
}

static int dev_rdf_tc(Ctx &c, Ctx::Lane &L, cudaStream_t st, const float *d_sel1, int n1, const float *d_sel2, int n2,
                      const int *d_gid, int nGroups, unsigned long long *d_counts, int nBins,
                      const float *d_cell, int nF, int fb, int fe, int sameSel, float dr,
                      const int *d_idx1, const int *d_idx2, double T)
{
    const bool grouped = d_gid != nullptr;
    const size_t base = tc_smem_base();
    const size_t budget = c.smemOptin > base ? c.smemOptin - base : 0;
    // The [groups][bins] histogram lives in shared memory.  When all groups do not fit (a protein of more than
    // ~230 residues at 149 bins) the pair kernel runs once per BATCH of groups over the same packed frames: atoms
    // of other groups are filtered like any other and dropped at the histogram (the reference makes one full
    // call per residue, dataAnalysis/RadialDensity.py:84-103).
    if ((size_t)nBins * 4 > budget)
        return fail(NDA_ERR_ARG, "outSize = %d histogram cells do not fit in shared memory (max %zu)", nBins, budget / 4);
    const int groupsPerBatch = (int)std::min<size_t>((size_t)nGroups, budget / ((size_t)nBins * 4));
    const size_t Hb = (size_t)groupsPerBatch * nBins;
    const int copies = (int)std::max<size_t>(1, std::min<size_t>(TC_EPI_WARPS, budget / (Hb * 4)));
    const size_t smem = base + (size_t)copies * Hb * 4;
    // sel1 is streamed in 128-atom blocks (TMEM lanes), sel2 is resident in 256-atom chunks (columns)
    const int n1Pad = round_up(n1, TC_SPS * TC_SBLK);          // whole two-block stages (= TC_RBLK)
    const int n2Pad = sameSel ? n1Pad : round_up(n2, TC_RBLK);
    const size_t bytesPerFrame = (size_t)n1Pad * (16 + 32) + (sameSel ? (size_t)n1Pad * 32 : (size_t)n2Pad * (16 + 32));
    const int pass = frames_per_pass(bytesPerFrame, fe - fb);

    CK(L.packA.reserve((size_t)n1Pad * pass * sizeof(float4)));
    if (!sameSel) CK(L.packB.reserve((size_t)n2Pad * pass * sizeof(float4)));
    CK(L.embA.reserve((size_t)n1Pad * pass * 32));
    CK(L.embB.reserve((size_t)n2Pad * pass * 32));
    CK(L.fp.reserve((size_t)pass * sizeof(FrameParams)));
    CK(L.thrDot.reserve((size_t)pass * sizeof(TcFrame)));
    CK(L.maxAbs.reserve((size_t)pass * sizeof(unsigned)));
    unsigned long long *stats = reinterpret_cast<unsigned long long *>(L.ctl.as<unsigned char>() + 16);

    RdfTcArgs a;
    TcGeom &g = a.g;
    g.n1 = n1; g.n1Pad = n1Pad; g.n2 = n2; g.n2Pad = n2Pad;
    g.nRChunks = (n2 + TC_RBLK - 1) / TC_RBLK;
    g.nSBlocks = n1Pad / TC_SBLK;                              // even; a trailing all-padding block has no candidates
    g.sameSel = sameSel ? 1 : 0;
    g.stats = stats;
    g.origInStage = 1;
    g.trace = nullptr;
    a.counts = d_counts; a.nBins = nBins; a.nGroups = groupsPerBatch; a.groupBase = 0; a.histCopies = copies;
    a.dr = dr;
    a.fastStrict = c.cfg.strictSlow ? 0 : 1;
    tc_chunking(g, c.numSMs, pass, 8);
    const void *kfn = grouped ? (const void *)rdf_tc_kernel<true> : (const void *)rdf_tc_kernel<false>;
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    for (int f0 = fb; f0 < fe; f0 += pass) {
        const int nFc = std::min(pass, fe - f0);
        CK(cudaMemsetAsync(L.maxAbs.p, 0, (size_t)nFc * sizeof(unsigned), st));
        // the A-role embedding needs no threshold: it is written by the pack kernel itself
        int rc = launch_pack(c, st, d_sel1, nF, d_idx1, n1, n1Pad, f0, nFc, sameSel ? d_gid : nullptr,
                             L.packA.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, nullptr, 0, L.embA.as<__half>());
        if (rc) return rc;
        if (!sameSel) {
            rc = launch_pack(c, st, d_sel2, nF, d_idx2, n2, n2Pad, f0, nFc, d_gid,
                             L.packB.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, nullptr, 0);
            if (rc) return rc;
        }
        g.A = L.packA.as<float4>();
        g.B = sameSel ? g.A : L.packB.as<float4>();
        prep_frames_tc_kernel<<<(nFc + 127) / 128, 128, 0, st>>>(d_cell, f0, nFc, L.maxAbs.as<unsigned>(), T, L.fp.as<FrameParams>(),
                                                                -1.0, c.cfg.tcRangeFactor * 1.2, L.thrDot.as<TcFrame>(), stats);
        LAUNCHED();
        rc = launch_embed(st, g.B, n2Pad, nFc, d_cell, f0, L.thrDot.as<TcFrame>(), L.embB.as<__half>(), 0);
        if (rc) return rc;
        g.EA = L.embA.as<__half>(); g.EB = L.embB.as<__half>();
        g.tf = L.thrDot.as<TcFrame>();
        const unsigned long long items = (unsigned long long)nFc * g.nRChunks * g.nSChunks;
        if (items > 0xfffffff0ull) return fail(NDA_ERR_ARG, "too many work items in one pass");
        g.nItems = (unsigned)items;
        const int grid = (int)std::min<unsigned long long>(items, (unsigned long long)c.numSMs);
        for (int g0 = 0; g0 < nGroups; g0 += groupsPerBatch) {
            a.groupBase = g0;
            a.nGroups = std::min(groupsPerBatch, nGroups - g0);
            a.counts = d_counts + (size_t)g0 * nBins;
            NvtxRange nv("nda:rdf_tc_kernel", c.cfg.nvtx);
            CK(c.kev_record(st));
            if (grouped) rdf_tc_kernel<true><<<grid, TC_THREADS, smem, st>>>(a);
            else         rdf_tc_kernel<false><<<grid, TC_THREADS, smem, st>>>(a);
            LAUNCHED();
            CK(c.kev_record(st));
        }
    }
    return NDA_OK;
}

static int dev_within_tc(Ctx &c, Ctx::Lane &L, cudaStream_t st, const float *d_all, int nF,
                         const int *d_refSel, int refSize, const int *d_outSel, int outSize,
                         unsigned *d_bits, const float *d_cell, float d2, double T, int fb, int fe)
{
    const int words = (outSize + 31) / 32;
    // outSel is streamed in 128-atom blocks (TMEM lanes), refSel is resident in 256-atom chunks (columns)
    const int nAPad = round_up(outSize, TC_SPS * TC_SBLK);     // whole two-block stages
    const int nBPad = round_up(refSize, TC_RBLK);
    const size_t bytesPerFrame = ((size_t)nAPad + nBPad) * (16 + 32);
    int pass = frames_per_pass(bytesPerFrame, fe - fb);
    // NDA_SUBPASS_FRAMES: frame sub-passes small enough that what pack / embed write is still in the 126 MB L2
    // when the pair kernel reads it, and is overwritten there by the next sub-pass before it reaches DRAM
    if (c.cfg.subpassFrames > 0) pass = std::min(pass, c.cfg.subpassFrames);
    CK(L.packA.reserve((size_t)nAPad * pass * sizeof(float4)));
    CK(L.packB.reserve((size_t)nBPad * pass * sizeof(float4)));
    CK(L.embA.reserve((size_t)nAPad * pass * 32));
    CK(L.embB.reserve((size_t)nBPad * pass * 32));
    CK(L.fp.reserve((size_t)pass * sizeof(FrameParams)));
    CK(L.thrDot.reserve((size_t)pass * sizeof(TcFrame)));
    CK(L.maxAbs.reserve((size_t)pass * sizeof(unsigned)));
    unsigned long long *stats = reinterpret_cast<unsigned long long *>(L.ctl.as<unsigned char>() + 16);

    WithinTcArgs a;
    TcGeom &g = a.g;
    g.n1 = outSize; g.n1Pad = nAPad; g.n2 = refSize; g.n2Pad = nBPad;
    g.nRChunks = (refSize + TC_RBLK - 1) / TC_RBLK;
    g.nSBlocks = nAPad / TC_SBLK;                              // even; a trailing all-padding block has no candidates
    g.sameSel = 0;
    g.stats = stats;
    g.origInStage = 0;
    g.trace = nullptr;
    if (c.cfg.tcTrace == 1) {                                  // debugging aid: clock64 trace of CTA 0 (tools/tc_trace.py)
        CK(c.sum.reserve(256 * 32 * 8));
        CK(cudaMemsetAsync(c.sum.p, 0, 256 * 32 * 8, st));
        g.trace = c.sum.as<long long>();
    }
    tc_chunking(g, c.numSMs, pass, 16, TC_WITHIN_MAX_STEPS);   // a warp records at most one (step, lanes) entry per step
    a.wordsPerFrame = words;
    a.d2 = d2;
    a.literal = g_within_pure.load() ? 0 : 1;
    const size_t smem = tc_smem_base() + TC_WITHIN_STAGE_BYTES;
    CK(cudaFuncSetAttribute((const void *)within_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));

    for (int f0 = fb; f0 < fe; f0 += pass) {
        const int nFc = std::min(pass, fe - f0);
        CK(cudaMemsetAsync(L.maxAbs.p, 0, (size_t)nFc * sizeof(unsigned), st));
        a.bits = d_bits + (size_t)(f0 - fb) * words;
        CK(cudaMemsetAsync(a.bits, 0, (size_t)nFc * words * sizeof(unsigned), st));   // found atoms are OR-ed in
        int rc = launch_pack(c, st, d_all, nF, d_outSel, outSize, nAPad, f0, nFc, nullptr,
                             L.packA.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, nullptr, 0, L.embA.as<__half>());
        if (rc) return rc;
        rc = launch_pack(c, st, d_all, nF, d_refSel, refSize, nBPad, f0, nFc, nullptr,
                         L.packB.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, nullptr, 0);
        if (rc) return rc;
        g.A = L.packA.as<float4>();
        g.B = L.packB.as<float4>();
        // the "sure" shortcut needs the pure predicate to be what the reference computes: with the literal
        // per-axis early-outs (getWithin.cpp:64-71) that holds for distance >= 1 (SURVEY B-2)
        const double sureD2 = (!a.literal || d2 >= 1.0f) && !c.cfg.tcNoSure ? (double)d2 : -1.0;
        prep_frames_tc_kernel<<<(nFc + 127) / 128, 128, 0, st>>>(d_cell, f0, nFc, L.maxAbs.as<unsigned>(), T, L.fp.as<FrameParams>(),
                                                                sureD2, c.cfg.tcRangeFactor * 1.2, L.thrDot.as<TcFrame>(), stats);
        LAUNCHED();
        rc = launch_embed(st, g.B, nBPad, nFc, d_cell, f0, L.thrDot.as<TcFrame>(), L.embB.as<__half>(), 0);
        if (rc) return rc;
        g.EA = L.embA.as<__half>(); g.EB = L.embB.as<__half>();
        g.tf = L.thrDot.as<TcFrame>();
        const unsigned long long items = (unsigned long long)nFc * g.nRChunks * g.nSChunks;
        if (items > 0xfffffff0ull) return fail(NDA_ERR_ARG, "too many work items in one pass");
        g.nItems = (unsigned)items;
        const int grid = (int)std::min<unsigned long long>(items, (unsigned long long)c.numSMs);
        {
            NvtxRange nv("nda:within_tc_kernel", c.cfg.nvtx);
            CK(c.kev_record(st));
            within_tc_kernel<<<grid, TC_THREADS, smem, st>>>(a);
            LAUNCHED();
            CK(c.kev_record(st));
        }
    }
    return NDA_OK;
}

static int dev_rdf(Ctx &c, Ctx::Lane &L, cudaStream_t st, const float *d_sel1, int n1, const float *d_sel2, int n2,
                   const int *d_gid, int nGroups, unsigned long long *d_counts, int nBins,
                   const float *d_cell, int nF, int fb, int fe, int sameSel, float dr,
                   const int *d_idx1 = nullptr, const int *d_idx2 = nullptr,   // optional row indices into d_sel1/2
                   const float *h_cell = nullptr)                              // host copy of d_cell, when the caller has one
{
    if (n1 <= 0 || n2 <= 0 || fe <= fb || nBins <= 0) return NDA_OK;
    if (nGroups < 1) nGroups = 1;
    const bool grouped = d_gid != nullptr;
    if (!grouped) nGroups = 1;
    if (c.cfg.filterTc) {
        const double lim0 = (double)nBins * fabs((double)dr);
        const double T0 = lim0 * lim0 * (1.0 + 9.5367431640625e-07);
        const float *hc = nullptr;
        int rc0 = host_cell_view(c, st, h_cell, d_cell, fb, fe, &hc);
        if (rc0) return rc0;
        const bool okTc = tc_frames_eligible(c.cfg, hc, fb, fe, T0, 1.0, n1, n2);
        if (c.cfg.debug) fprintf(stderr, "[nda] rdf: tensor filter %s (T %.3f, frames %d..%d, cell0 %.2f %.2f %.2f)\n",
                                         okTc ? "on" : "off", T0, fb, fe, hc[3 * (size_t)fb], hc[3 * (size_t)fb + 1], hc[3 * (size_t)fb + 2]);
        if (okTc) {
            g_last_filter.store(100);
            return dev_rdf_tc(c, L, st, d_sel1, n1, d_sel2, n2, d_gid, nGroups, d_counts, nBins, d_cell, nF, fb, fe,
                              sameSel, dr, d_idx1, d_idx2, T0);
        }
    }
    // shared memory: as many warp-private histogram copies as still allow PAIR_MIN_CTAS CTAs per SM
    const size_t perCta = ((size_t)227 * 1024) / PAIR_MIN_CTAS - 1024;
    // The kernel with 256-atom B tiles (the grouped one; an absent group array means group 0 for every atom) also
    // serves SMALL ungrouped calls: with 1024-atom tiles a call like BASELINE config 1 (1000 atoms x 100 frames) has 200
    // work items for ~590 resident CTAs and runs on a third of the GPU.
    const long long items1024 = (long long)(fe - fb) * ((n1 + PAIR_ATILE - 1) / PAIR_ATILE) * ((n2 + PAIR_BTILE - 1) / PAIR_BTILE);
    const bool smallB = grouped || items1024 < 16LL * c.numSMs;
    const int bTile = smallB ? PAIR_BTILE_GROUPED : PAIR_BTILE;
    const size_t base = pair_smem_base(bTile);
    size_t histBudget = (perCta > base) ? perCta - base : 0;
    // groups in batches that fit the shared histogram (see dev_rdf_tc)
    const size_t maxCells = (c.smemOptin > base ? c.smemOptin - base : 0) / 4;
    if ((size_t)nBins > maxCells)
        return fail(NDA_ERR_ARG, "outSize = %d histogram cells do not fit in shared memory (max %zu)", nBins, maxCells);
    const int groupsPerBatch = (int)std::min<size_t>((size_t)nGroups, maxCells / (size_t)nBins);
    const size_t Hb = (size_t)groupsPerBatch * nBins;
    int copies = (int)std::min<size_t>(PAIR_WARPS, histBudget / (Hb * 4));
    if (copies < 1) copies = 1;
    const size_t smem = base + (size_t)copies * Hb * 4;
    const int n1Pad = sameSel ? round_up(n1, PAIR_BTILE) : round_up(n1, PAIR_ATILE);
    const int n2Pad = sameSel ? n1Pad : round_up(n2, PAIR_BTILE);
    const int filter = c.cfg.filterKind;
    g_last_filter.store(filter);
    const bool fold = filter != kFilterRint;                 // wrapped copies needed
    const int pairLayout = is_fold2(filter) ? 1 : 0;
    const size_t bytesPerFrame = ((size_t)n1Pad + (sameSel ? 0 : (size_t)n2Pad)) * sizeof(float4) * (fold ? 2 : 1);
    const int pass = frames_per_pass(bytesPerFrame, fe - fb);

    CK(L.packA.reserve((size_t)n1Pad * pass * sizeof(float4)));
    if (!sameSel) CK(L.packB.reserve((size_t)n2Pad * pass * sizeof(float4)));
    if (fold) {
        CK(L.packAw.reserve((size_t)n1Pad * pass * sizeof(float4)));
        if (!sameSel) CK(L.packBw.reserve((size_t)n2Pad * pass * sizeof(float4)));
    }
    CK(L.fp.reserve((size_t)pass * sizeof(FrameParams)));
    CK(L.maxAbs.reserve((size_t)pass * sizeof(unsigned)));

    const double drd = (double)dr;
    const double lim = (double)nBins * fabs(drd);
    const double T = lim * lim * (1.0 + 9.5367431640625e-07);       // (nBins*dr)^2 (1 + 2^-20)

    unsigned *work = L.ctl.as<unsigned>();
    unsigned long long *stats = reinterpret_cast<unsigned long long *>(L.ctl.as<unsigned char>() + 16);

    RdfArgs a;
    a.n1 = n1; a.n1Pad = n1Pad; a.n2 = n2; a.n2Pad = n2Pad;
    a.nATiles = (n1 + PAIR_ATILE - 1) / PAIR_ATILE;
    a.nBTiles = (n2 + bTile - 1) / bTile;
    a.counts = d_counts; a.nBins = nBins; a.nGroups = groupsPerBatch; a.groupBase = 0; a.histCopies = copies;
    a.sameSel = sameSel ? 1 : 0; a.dr = dr;
    a.fastStrict = c.cfg.strictSlow ? 0 : 1;
    a.workCounter = work; a.stats = stats;

    int ctasPerSM = 0;
    const void *kfn = nullptr;
#define NDA_RDF_CASE(G, F) case F: kfn = (const void *)rdf_kernel<G, F>; break;
    if (smallB)  switch (filter) { NDA_RDF_CASE(true, 0) NDA_RDF_CASE(true, 1) NDA_RDF_CASE(true, 2) NDA_RDF_CASE(true, 3) NDA_RDF_CASE(true, 4) NDA_RDF_CASE(true, 5) }
    else         switch (filter) { NDA_RDF_CASE(false, 0) NDA_RDF_CASE(false, 1) NDA_RDF_CASE(false, 2) NDA_RDF_CASE(false, 3) NDA_RDF_CASE(false, 4) NDA_RDF_CASE(false, 5) }
#undef NDA_RDF_CASE
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSM, kfn, PAIR_THREADS, smem));
    ctasPerSM = tuned_ctas(c.cfg, ctasPerSM);
    // B chunk per work item: aim at >= 16 items per resident CTA so the dynamic scheduler can
    // balance the triangular (sameSel) workload, but keep an item >= 2 tiles (~1e6 pairs)
    {
        const long long target = 16LL * ctasPerSM * c.numSMs;
        const long long rows = (long long)pass * a.nATiles;
        const long long chunksWanted = std::max<long long>(1, (target + rows - 1) / rows);
        long long ct = (a.nBTiles + chunksWanted - 1) / chunksWanted;
        ct = std::max<long long>((smallB && !grouped) ? 1 : 2, std::min<long long>(32, ct));
        a.chunkTiles = (int)std::min<long long>(ct, a.nBTiles);
        a.nBChunks = (a.nBTiles + a.chunkTiles - 1) / a.chunkTiles;
    }

    for (int f0 = fb; f0 < fe; f0 += pass) {
        const int nFc = std::min(pass, fe - f0);
        CK(cudaMemsetAsync(L.maxAbs.p, 0, (size_t)nFc * sizeof(unsigned), st));
        CK(cudaMemsetAsync(work, 0, 8, st));
        int rc = launch_pack(c, st, d_sel1, nF, d_idx1, n1, n1Pad, f0, nFc, sameSel ? d_gid : nullptr,
                             L.packA.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, fold ? L.packAw.as<float4>() : nullptr, pairLayout);
        if (rc) return rc;
        if (!sameSel) {
            rc = launch_pack(c, st, d_sel2, nF, d_idx2, n2, n2Pad, f0, nFc, d_gid,
                             L.packB.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, fold ? L.packBw.as<float4>() : nullptr, pairLayout);
            if (rc) return rc;
        }
        prep_frames_kernel<<<(nFc + 127) / 128, 128, 0, st>>>(d_cell, f0, nFc, L.maxAbs.as<unsigned>(), T,
                                                             L.fp.as<FrameParams>(), stats, c.cfg.denseFraction);
        LAUNCHED();
        a.A = L.packA.as<float4>();
        a.B = sameSel ? a.A : L.packB.as<float4>();
        a.Aw = fold ? L.packAw.as<float4>() : a.A;
        a.Bw = fold ? (sameSel ? a.Aw : L.packBw.as<float4>()) : a.B;
        a.fp = L.fp.as<FrameParams>();
        const unsigned long long items = (unsigned long long)nFc * a.nATiles * a.nBChunks;
        if (items > 0xfffffff0ull) return fail(NDA_ERR_ARG, "too many work items in one pass");
        a.nItems = (unsigned)items;
        const int grid = (int)std::min<unsigned long long>(items, (unsigned long long)ctasPerSM * c.numSMs);
#define NDA_RDF_LAUNCH(G, F) case F: rdf_kernel<G, F><<<grid, PAIR_THREADS, smem, st>>>(a); break;
        for (int g0 = 0; g0 < nGroups; g0 += groupsPerBatch) {
            a.groupBase = g0;
            a.nGroups = std::min(groupsPerBatch, nGroups - g0);
            a.counts = d_counts + (size_t)g0 * nBins;
            if (g0 > 0) CK(cudaMemsetAsync(work, 0, 8, st));
            NvtxRange nv("nda:rdf_kernel", c.cfg.nvtx);
            CK(c.kev_record(st));
            if (smallB)  switch (filter) { NDA_RDF_LAUNCH(true, 0) NDA_RDF_LAUNCH(true, 1) NDA_RDF_LAUNCH(true, 2) NDA_RDF_LAUNCH(true, 3) NDA_RDF_LAUNCH(true, 4) NDA_RDF_LAUNCH(true, 5) }
            else         switch (filter) { NDA_RDF_LAUNCH(false, 0) NDA_RDF_LAUNCH(false, 1) NDA_RDF_LAUNCH(false, 2) NDA_RDF_LAUNCH(false, 3) NDA_RDF_LAUNCH(false, 4) NDA_RDF_LAUNCH(false, 5) }
            LAUNCHED();
            CK(c.kev_record(st));
        }
#undef NDA_RDF_LAUNCH
    }
    return NDA_OK;
}

static int dev_within(Ctx &c, Ctx::Lane &L, cudaStream_t st, const float *d_all, int nAtoms, int nF,
                      const int *d_refSel, int refSize, const int *d_outSel, int outSize,
                      unsigned *d_bits, int *d_out, const float *d_cell, float distance, int fb, int fe,
                      const float *h_cell = nullptr)
{
    (void)nAtoms;
    if (outSize <= 0 || fe <= fb) return NDA_OK;
    const int words = (outSize + 31) / 32;
    bool doneTc = false;
    if (refSize > 0 && c.cfg.filterTc) {
        const float d2t = distance * distance;
        const double Tt = (double)d2t * (1.0 + 9.5367431640625e-07);
        const float *hc = nullptr;
        int rc0 = host_cell_view(c, st, h_cell, d_cell, fb, fe, &hc);
        if (rc0) return rc0;
        if (tc_frames_eligible(c.cfg, hc, fb, fe, Tt, 1.0, outSize, refSize)) {
            rc0 = dev_within_tc(c, L, st, d_all, nF, d_refSel, refSize, d_outSel, outSize, d_bits, d_cell, d2t, Tt, fb, fe);
            if (rc0) return rc0;
            doneTc = true;
            g_last_filter.store(100);
        }
    }
    if (doneTc) {
    } else if (refSize <= 0) {
        CK(cudaMemsetAsync(d_bits, 0, (size_t)(fe - fb) * words * sizeof(unsigned), st));
    } else {
        const int nAPad = round_up(outSize, PAIR_ATILE);
        const int nBPad = round_up(refSize, PAIR_BTILE);
        const int filter = c.cfg.filterKind;
        g_last_filter.store(filter);
        const bool fold = filter != kFilterRint;             // wrapped copies needed
        const int pairLayout = is_fold2(filter) ? 1 : 0;
        const size_t bytesPerFrame = ((size_t)nAPad + nBPad) * sizeof(float4) * (fold ? 2 : 1);
        const int pass = frames_per_pass(bytesPerFrame, fe - fb);
        CK(L.packA.reserve((size_t)nAPad * pass * sizeof(float4)));
        CK(L.packB.reserve((size_t)nBPad * pass * sizeof(float4)));
        if (fold) {
            CK(L.packAw.reserve((size_t)nAPad * pass * sizeof(float4)));
            CK(L.packBw.reserve((size_t)nBPad * pass * sizeof(float4)));
        }
        CK(L.fp.reserve((size_t)pass * sizeof(FrameParams)));
        CK(L.maxAbs.reserve((size_t)pass * sizeof(unsigned)));
        unsigned *work = L.ctl.as<unsigned>();
        unsigned long long *stats = reinterpret_cast<unsigned long long *>(L.ctl.as<unsigned char>() + 16);

        const float d2 = distance * distance;                  // one float32 rounding, getWithin.cpp:29
        const double T = (double)d2 * (1.0 + 9.5367431640625e-07);

        WithinArgs a;
        a.nA = outSize; a.nAPad = nAPad; a.nB = refSize; a.nBPad = nBPad;
        a.nATiles = nAPad / PAIR_ATILE;
        a.nBTiles = (refSize + PAIR_BTILE - 1) / PAIR_BTILE;
        a.wordsPerFrame = words;
        a.d2 = d2;
        a.literal = g_within_pure.load() ? 0 : 1;
        a.workCounter = work; a.stats = stats;
        const size_t smem = PAIR_SMEM_BASE;
        const void *kfn = nullptr;
        switch (filter) {
        case 0: kfn = (const void *)within_kernel<0>; break;
        case 1: kfn = (const void *)within_kernel<1>; break;
        case 2: kfn = (const void *)within_kernel<2>; break;
        case 3: kfn = (const void *)within_kernel<3>; break;
        case 4: kfn = (const void *)within_kernel<4>; break;
        default: kfn = (const void *)within_kernel<5>; break;
        }
        CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int ctasPerSM = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctasPerSM, kfn, PAIR_THREADS, smem));
        ctasPerSM = tuned_ctas(c.cfg, ctasPerSM);

        for (int f0 = fb; f0 < fe; f0 += pass) {
            const int nFc = std::min(pass, fe - f0);
            CK(cudaMemsetAsync(L.maxAbs.p, 0, (size_t)nFc * sizeof(unsigned), st));
            CK(cudaMemsetAsync(work, 0, 8, st));
            int rc = launch_pack(c, st, d_all, nF, d_outSel, outSize, nAPad, f0, nFc, nullptr,
                                 L.packA.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, fold ? L.packAw.as<float4>() : nullptr, pairLayout);
            if (rc) return rc;
            rc = launch_pack(c, st, d_all, nF, d_refSel, refSize, nBPad, f0, nFc, nullptr,
                             L.packB.as<float4>(), L.maxAbs.as<unsigned>(), d_cell, fold ? L.packBw.as<float4>() : nullptr, pairLayout);
            if (rc) return rc;
            prep_frames_kernel<<<(nFc + 127) / 128, 128, 0, st>>>(d_cell, f0, nFc, L.maxAbs.as<unsigned>(), T,
                                                                 L.fp.as<FrameParams>(), stats, 0.0);
            LAUNCHED();
            a.A = L.packA.as<float4>();
            a.B = L.packB.as<float4>();
            a.Aw = fold ? L.packAw.as<float4>() : a.A;
            a.Bw = fold ? L.packBw.as<float4>() : a.B;
            a.fp = L.fp.as<FrameParams>();
            a.bits = d_bits + (size_t)(f0 - fb) * words;
            const unsigned long long items = (unsigned long long)nFc * a.nATiles;
            if (items > 0xfffffff0ull) return fail(NDA_ERR_ARG, "too many work items in one pass");
            a.nItems = (unsigned)items;
            const int grid = (int)std::min<unsigned long long>(items, (unsigned long long)ctasPerSM * c.numSMs);
            CK(c.kev_record(st));
            switch (filter) {
            case 0: within_kernel<0><<<grid, PAIR_THREADS, smem, st>>>(a); break;
            case 1: within_kernel<1><<<grid, PAIR_THREADS, smem, st>>>(a); break;
            case 2: within_kernel<2><<<grid, PAIR_THREADS, smem, st>>>(a); break;
            case 3: within_kernel<3><<<grid, PAIR_THREADS, smem, st>>>(a); break;
            case 4: within_kernel<4><<<grid, PAIR_THREADS, smem, st>>>(a); break;
            default: within_kernel<5><<<grid, PAIR_THREADS, smem, st>>>(a); break;
            }
            LAUNCHED();
            CK(c.kev_record(st));
        }
    }
    if (d_out) {
        dim3 grid((fe - fb + 127) / 128, std::min(words, 65535));
        expand_mask_kernel<<<grid, 128, 0, st>>>(d_bits, words, d_outSel, outSize, nF, fb, fe - fb, d_out);
        LAUNCHED();
    }
    return NDA_OK;
}

static int dev_distances(Ctx &c, cudaStream_t st, const float *d_sel1, int n1, const float *d_sel2, int n2,
                         double *d_sum, const float *d_cell, int nF, int fb, int fe, int sameSel,
                         const int *d_idx1 = nullptr, const int *d_idx2 = nullptr)
{
    if (n1 <= 0 || n2 <= 0 || fe <= fb) return NDA_OK;
    const int zt = (n1 + DIST_NA - 1) / DIST_NA;
    if (zt > 65535) return fail(NDA_ERR_ARG, "selection too large for one launch");
    const void *kfn = (const void *)distances_kernel;
    const size_t smem = sizeof(DistSmem);
    CK(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kfn, DIST_JT, smem));
    const long long slots = (long long)std::max(occ, 1) * c.numSMs;
    // frames per work item: ONE wave of equal items (every resident block gets exactly one, so nothing waits for a
    // straggler), each as long as that allows -- the pipeline prologue and the atomics are per item
    const long long jTiles = (n2 + DIST_JT - 1) / DIST_JT;
    const long long wantChunks = std::max<long long>(1, slots / (jTiles * zt));
    long long stagesLL = (((long long)(fe - fb) + wantChunks - 1) / wantChunks + DIST_FS - 1) / DIST_FS;
    const int stages = (int)std::max<long long>(1, std::min<long long>(stagesLL, 1 << 20));
    const long long nFChunks = ((long long)(fe - fb) + (long long)stages * DIST_FS - 1) / ((long long)stages * DIST_FS);
    const long long items = jTiles * nFChunks;
    if (items > 0xfffffff0ll) return fail(NDA_ERR_ARG, "too many work items in one launch");
    dim3 grid((unsigned)std::min<long long>(items, std::max<long long>(1, slots / zt)), zt);
    CK(c.kev_record(st));
    distances_kernel<<<grid, DIST_JT, smem, st>>>(d_sel1, n1, d_sel2, n2, d_cell, nF, fb, fe, sameSel ? 1 : 0, d_sum,
                                                  d_idx1, d_idx2, stages, (int)nFChunks, (unsigned)items, -0.0f);
    LAUNCHED();
    CK(c.kev_record(st));
    return NDA_OK;
}

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
// Pipelined upload of frame chunks of reference-layout arrays ([row][nF][3] float32).
//
//  * A row (one atom, all frames) is contiguous, so a run of rows with a constant index stride and a
//    frame window is ONE strided 2-D copy (water oxygens at 1231 + 3k: one copy).
//  * Pinned (or registered) host memory is DMA'd directly.  Pageable memory -- what numpy gives the
//    reference's callers -- is first gathered by a few host threads into pinned staging and then
//    DMA'd as one contiguous block; the driver's own pageable path moves narrow rows one by one.
//  * Two device buffers per array alternate between the copy stream and the compute stream:
//    the upload of chunk k+1 overlaps pack + kernel of chunk k.
static void runs_from_rows(const std::vector<int> &rows, std::vector<RowRun> &runs)
{
    runs.clear();
    size_t i = 0;
    while (i < rows.size()) {
        size_t j = i + 1;
        int stride = 1;
        if (j < rows.size()) {
            stride = rows[j] - rows[i];
            while (j + 1 < rows.size() && rows[j + 1] - rows[j] == stride) ++j;
            ++j;
        }
        runs.push_back({(int)i, rows[i], stride, (int)(j - i)});
        i = j;
    }
}

// rows needed by a within call and the selections re-expressed in uploaded-row numbers
static void plan_rows(const int *refSel, int refSize, const int *outSel, int outSize, int nAtoms,
                      std::vector<int> &remap, std::vector<RowRun> &runs, int &nRows)
{
    std::vector<unsigned char> used((size_t)nAtoms, 0);
    for (int i = 0; i < refSize; ++i) used[refSel[i]] = 1;
    for (int i = 0; i < outSize; ++i) used[outSel[i]] = 1;
    std::vector<int> rows;
    for (int a = 0; a < nAtoms; ++a) if (used[a]) rows.push_back(a);
    runs_from_rows(rows, runs);
    if (runs.size() > 512 && !rows.empty()) {              // scattered selection: copy the covering range
        const int lo = rows.front(), hi = rows.back();
        rows.clear();
        for (int a = lo; a <= hi; ++a) rows.push_back(a);
        runs.assign(1, RowRun{0, lo, 1, hi - lo + 1});
    }
    remap.assign((size_t)nAtoms, -1);
    for (size_t r = 0; r < rows.size(); ++r) remap[rows[r]] = (int)r;
    nRows = (int)rows.size();
}

static bool host_ptr_is_pinned(const void *p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// The small per-call arrays (cell edges, selections, group ids) go host -> device on the COPY stream,
// ahead of the first bulk chunk, from our own pinned staging.  On the compute stream they would sit in
// another channel of the same copy engine: when the caller's array is pinned (a true async DMA) the
// engine keeps serving the copy stream's channel while it has ready work, and the 12 KB of cell edges
// waited for TWO bulk chunks (~0.9 ms) -- with it the first chunk's kernels (seen in bench.py).
struct SetupItem { void *dst; const void *src; size_t bytes; };

static int upload_setup(Ctx &c, const SetupItem *items, int n)
{
    size_t total = 0;
    for (int i = 0; i < n; ++i) total += (items[i].bytes + 255) & ~(size_t)255;
    if (total > c.setupCap) {
        if (c.setupStage) { cudaFreeHost(c.setupStage); c.setupStage = nullptr; c.setupCap = 0; }
        CK(cudaMallocHost(&c.setupStage, total + total / 2 + 4096));
        c.setupCap = total + total / 2 + 4096;
    }
    size_t off = 0;
    for (int i = 0; i < n; ++i) {
        if (!items[i].bytes) continue;
        char *h = static_cast<char *>(c.setupStage) + off;
        memcpy(h, items[i].src, items[i].bytes);
        CK(cudaMemcpyAsync(items[i].dst, h, items[i].bytes, cudaMemcpyHostToDevice, c.copyStream));
        off += (items[i].bytes + 255) & ~(size_t)255;
    }
    CK(cudaEventRecord(c.evSetupCopy, c.copyStream));
    return NDA_OK;
}

struct UploadArray {
    const float *host = nullptr;
    int nRows = 0;                       // rows uploaded (after the row plan)
    std::vector<RowRun> runs;
    bool pinned = false;
    size_t stageOff = 0;                 // byte offset of this array inside a staging buffer
};

struct Uploader {
    Ctx &c;
    int nF, fb, fe, chunk = 0, nChunks = 0;     // chunk = the largest chunk (what the buffers are sized for)
    std::vector<int> bnd;                        // chunk k covers frames [bnd[k], bnd[k+1])
    std::vector<UploadArray> arr;
    bool taperTail = true;
    Uploader(Ctx &ctx, int nF_, int fb_, int fe_) : c(ctx), nF(nF_), fb(fb_), fe(fe_) {}

    void add(const float *host, int nRows)                              // all rows 0..nRows-1
    {
        UploadArray a;
        a.host = host; a.nRows = nRows;
        if (nRows > 0) a.runs.push_back({0, 0, 1, nRows});
        arr.push_back(a);
    }
    void add(const float *host, int nRows, const std::vector<RowRun> &runs)
    {
        UploadArray a;
        a.host = host; a.nRows = nRows; a.runs = runs;
        arr.push_back(a);
    }

    // choose the chunking, reserve device / staging buffers, mark both buffers free
    int begin(int maxChunks, int chunkOverride)
    {
        const int nFc = fe - fb;
        size_t bytesPerFrame = 0;
        for (auto &a : arr) { a.pinned = host_ptr_is_pinned(a.host); bytesPerFrame += (size_t)a.nRows * 12; }
        if (bytesPerFrame == 0) bytesPerFrame = 12;
        // ~20 MB per chunk: with the tensor-core kernels a chunk's compute is a fraction of its copy, and
        // every chunk costs ~30 us of copy set-up (measured: 110 MB in 5-6 chunks 2.6 ms, in 16 chunks 3.0 ms)
        nChunks = (int)std::min<size_t>((size_t)maxChunks, std::max<size_t>(1, bytesPerFrame * nFc / ((size_t)20 << 20)));
        if (chunkOverride >= 1) nChunks = chunkOverride;
        nChunks = std::max(1, std::min(nChunks, nFc));
        chunk = (nFc + nChunks - 1) / nChunks;
        const int maxChunk = (int)std::max<size_t>(1, ((size_t)3 << 30) / bytesPerFrame);   // <= 3 GB per buffer
        chunk = std::max(1, std::min(chunk, maxChunk));
        nChunks = (nFc + chunk - 1) / chunk;
        bnd.assign((size_t)nChunks + 1, fe);
        for (int k = 0; k < nChunks; ++k) bnd[k] = fb + k * chunk;
        // The copies are the bottleneck, so what the caller waits for after the last byte has arrived is the
        // compute (and the host post-processing) of the LAST chunk: taper the tail -- one more chunk, the last
        // two of 1/2 and 1/4 of the regular size (NDA_NO_TAPER=1 keeps them uniform).
        const bool noTaper = c.cfg.noTaper != 0;
        const int taperLevels = c.cfg.taperLevels;
        if (taperTail && !noTaper && nChunks >= 3 && nChunks < maxChunks + 1 && chunk * 4 <= maxChunk * 3 && nFc >= 16 * nChunks) {
            // chunk weights in units of a regular chunk: `levels` tail chunks of 1/2, 1/4, ... (the last one takes what
            // is left).  NDA_HEAD_TAPER=1 adds a HEAD of 1/4 and 1/2 for pageable inputs, so that the first DMA starts after
            // a shorter gather -- measured on a B200 host (config 2 from plain numpy arrays): 3.9-4.1 ms per call against
            // 3.1-3.2 ms without, because a narrow frame window means 540-byte row pieces 12 KB apart, which the host threads
            // gather far less efficiently than 2.5 KB pieces.  Off by default, kept as a switch for the record.
            bool staged = false;
            for (auto &a : arr) staged = staged || (!a.pinned && a.nRows > 0);
            std::vector<double> wts;
            if (staged && c.cfg.headTaper) { wts.push_back(0.25); wts.push_back(0.5); }
            const int levels = std::min(taperLevels, nChunks - 1);
            const int nReg = nChunks - 1;
            for (int k = 0; k < nReg; ++k) wts.push_back(1.0);
            double w = 0.5;
            for (int l = 0; l < levels; ++l, w *= 0.5) wts.push_back(w);
            double sum = 0.0;
            for (double x : wts) sum += x;
            const int n = (int)wts.size();
            bnd.assign((size_t)n + 1, fe);
            bnd[0] = fb;
            int big = 0;
            double upTo = 0.0;
            for (int k = 1; k < n; ++k) {
                upTo += wts[k - 1] * (double)nFc / sum;
                bnd[k] = std::min(fe - (n - k), std::max(bnd[k - 1] + 1, fb + (int)(upTo + 0.5)));
            }
            for (int k = 0; k < n; ++k) big = std::max(big, bnd[k + 1] - bnd[k]);
            nChunks = n;
            chunk = big;
        }
        size_t stageBytes = 0;
        for (size_t i = 0; i < arr.size(); ++i) {
            if (arr.size() > 2) return fail(NDA_ERR_ARG, "internal: at most two arrays per upload");
            CK(c.raw[i][0].reserve((size_t)arr[i].nRows * chunk * 12 + 16));
            if (nChunks > 1) CK(c.raw[i][1].reserve((size_t)arr[i].nRows * chunk * 12 + 16));
            if (!arr[i].pinned) { arr[i].stageOff = stageBytes; stageBytes += ((size_t)arr[i].nRows * chunk * 12 + 255) & ~(size_t)255; }
        }
        if (stageBytes) {
            CK(c.stage_reserve(0, stageBytes));
            if (nChunks > 1) CK(c.stage_reserve(1, stageBytes));
        }
        CK(cudaEventRecord(c.evDone[0], c.stream));
        CK(cudaEventRecord(c.evDone[1], c.stream));
        CK(cudaEventRecord(c.evCopied[0], c.copyStream));
        CK(cudaEventRecord(c.evCopied[1], c.copyStream));
        return NDA_OK;
    }

    int f0(int k) const { return bnd[k]; }
    int f1(int k) const { return bnd[k + 1]; }
    float *dev(int i, int k) const { return c.raw[i][k & 1].as<float>(); }

    // enqueue the upload of chunk k on the copy stream (gathering on the host first if pageable)
    // `side`, if given, is host work that does not depend on this chunk (the within path hands in the mask scatter of an
    // earlier chunk).  By default the whole team runs it after the gather.  NDA_SIDE_TEAM=1 runs it BESIDE the gather on a
    // quarter of the team -- measured on a 16-core B200 host (config 2 from plain numpy): 4.05-4.27 ms per call against
    // 3.92-4.06 ms, the gather loses as much with 12 threads as the overlap wins; kept as a switch for the record.
    // side(part, parts) must cover part = 0 .. parts - 1.
    int issue(int k, const std::function<void(int, int)> *side = nullptr)
    {
        NvtxRange nv("nda:upload_chunk", c.cfg.nvtx);
        const int b = k & 1, a0 = f0(k), wc = f1(k) - a0;
        const size_t w = (size_t)wc * 12;
        bool anyStage = false;
        for (auto &a : arr) anyStage = anyStage || (!a.pinned && a.nRows > 0);
        if (anyStage) CK(cudaEventSynchronize(c.evCopied[b]));           // staging buffer b has been consumed
        bool sideDone = side == nullptr;
        for (size_t i = 0; i < arr.size(); ++i) {
            UploadArray &a = arr[i];
            if (a.pinned || a.nRows == 0) continue;
            char *dst = static_cast<char *>(c.stage[b]) + a.stageOff;
            const float *src = a.host;
            const int nRuns = (int)a.runs.size();
            const long long nFl = nF;
            const int nRows = a.nRows;
            const bool par = (size_t)nRows * w > (1u << 20);
            const int T = par ? std::max(1, c.hostThreads) : 1;
            const int S = (!sideDone && c.cfg.sideTeam && T >= 8) ? T / 4 : 0;   // threads that run the side job
            #pragma omp parallel num_threads(T) if (par)
            {
                const int tid = omp_get_thread_num(), nt = omp_get_num_threads();
                const int s = (nt == T) ? S : 0;                              // a smaller team than asked for: no split
                if (tid >= nt - s) {
                    (*side)(tid - (nt - s), s);
                } else {
                    const int g = nt - s;
                    const int r0 = (int)((long long)nRows * tid / g), r1 = (int)((long long)nRows * (tid + 1) / g);
                    int q = 0;
                    for (int r = r0; r < r1; ++r) {
                        // locate the run of uploaded row r (runs are few and r ascends: the scan resumes)
                        while (q + 1 < nRuns && a.runs[q + 1].dstRow <= r) ++q;
                        const long long srcRow = a.runs[q].srcRow + (long long)(r - a.runs[q].dstRow) * a.runs[q].stride;
                        // (streaming stores were tried here: no faster for 2.5 KB rows, far slower for 24-byte rows)
                        memcpy(dst + (size_t)r * w, src + (srcRow * nFl + a0) * 3, w);
                    }
                }
                if (tid == 0 && s > 0) sideDone = true;                       // (every thread of the side team has been started)
            }
        }
        if (!sideDone) {                                                      // no gather to hide behind: the whole team
            const int T = std::max(1, c.hostThreads);
            #pragma omp parallel num_threads(T)
            (*side)(omp_get_thread_num(), omp_get_num_threads());
        }
        CK(cudaStreamWaitEvent(c.copyStream, c.evDone[b], 0));             // device buffer b is free again
        // strided 2-D DMA from pinned memory runs at about half the link rate on one copy engine,
        // so the rows of every run are split over two copy streams (NDA_COPY_STREAMS=1 disables)
        // (rows of >= 2 KB reach the link rate on one engine; NDA_COPY_STREAMS=1|2 overrides)
        const int csEnv = c.cfg.copyStreams;
        const bool twoStreams = csEnv == 2 || (csEnv != 1 && w < 2048);
        bool usedSecond = false;
        for (size_t i = 0; i < arr.size(); ++i) {
            UploadArray &a = arr[i];
            if (a.nRows == 0) continue;
            float *raw = dev((int)i, k);
            if (!a.pinned) {
                CK(cudaMemcpyAsync(raw, static_cast<char *>(c.stage[b]) + a.stageOff, (size_t)a.nRows * w,
                                   cudaMemcpyHostToDevice, c.copyStream));
                continue;
            }
            for (const RowRun &r0 : a.runs) {
                RowRun part[2] = {r0, r0};
                int nPart = 1;
                if (twoStreams && r0.count >= 64) {
                    part[0].count = r0.count / 2;
                    part[1].dstRow = r0.dstRow + part[0].count;
                    part[1].srcRow = r0.srcRow + part[0].count * r0.stride;
                    part[1].count = r0.count - part[0].count;
                    nPart = 2;
                }
                for (int h = 0; h < nPart; ++h) {
                    const RowRun &r = part[h];
                    cudaStream_t cs = h ? c.copyStream2 : c.copyStream;
                    if (h && !usedSecond) { CK(cudaStreamWaitEvent(c.copyStream2, c.evDone[b], 0)); usedSecond = true; }
                    const float *src = a.host + ((size_t)r.srcRow * nF + a0) * 3;
                    float *dst = raw + (size_t)r.dstRow * wc * 3;
                    if (r.count == 1 || (r.stride == 1 && wc == nF)) {
                        CK(cudaMemcpyAsync(dst, src, w * r.count, cudaMemcpyHostToDevice, cs));
                    } else if ((size_t)r.stride * nF * 12 > ((size_t)1 << 30)) {
                        for (int q = 0; q < r.count; ++q)          // pitch beyond the 2-D copy limit: row by row
                            CK(cudaMemcpyAsync(dst + (size_t)q * wc * 3, src + (size_t)q * r.stride * nF * 3, w,
                                               cudaMemcpyHostToDevice, cs));
                    } else {
                        CK(cudaMemcpy2DAsync(dst, w, src, (size_t)r.stride * nF * 12, w, (size_t)r.count,
                                             cudaMemcpyHostToDevice, cs));
                    }
                }
            }
        }
        if (usedSecond) {
            CK(cudaEventRecord(c.evJoin, c.copyStream2));
            CK(cudaStreamWaitEvent(c.copyStream, c.evJoin, 0));
        }
        CK(cudaEventRecord(c.evCopied[b], c.copyStream));
        return NDA_OK;
    }

    // compute stream: wait for chunk k's data / declare its buffers consumed
    int acquire(int k, cudaStream_t st) { CK(cudaStreamWaitEvent(st, c.evCopied[k & 1], 0)); return NDA_OK; }
    int release(int k, cudaStream_t st) { CK(cudaEventRecord(c.evDone[k & 1], st)); return NDA_OK; }
};

static cudaError_t reset_stats(Ctx &c, cudaStream_t st)
{
    c.statsPending = false;
    cudaError_t e = cudaMemsetAsync(c.lane[0].ctl.as<unsigned char>() + 16, 0, 16, st);
    if (e != cudaSuccess) return e;
    return cudaMemsetAsync(c.lane[1].ctl.as<unsigned char>() + 16, 0, 16, st);
}

static int read_stats(Ctx &c, double pairs, float ms)
{
    c.stats[0] = pairs;
    c.stats[3] = (double)ms;
    c.statsPending = true;               // two blocking 16-byte reads (~35 us): only when somebody asks
    return NDA_OK;
}

static int fetch_stats(Ctx &c)
{
    if (!c.statsPending) return NDA_OK;
    CK(cudaDeviceSynchronize());         // a device-pointer call may still be running on the caller's stream
    unsigned long long h[2][2] = {{0, 0}, {0, 0}};
    CK(cudaMemcpy(h[0], c.lane[0].ctl.as<unsigned char>() + 16, 16, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(h[1], c.lane[1].ctl.as<unsigned char>() + 16, 16, cudaMemcpyDeviceToHost));
    c.stats[1] = (double)(h[0][0] + h[1][0]);
    c.stats[2] = (double)(h[0][1] + h[1][1]);
    c.statsPending = false;
    return NDA_OK;
}

// host-side phase clock of one entry-point call (NDA_DEBUG=2 prints it)
struct HostTrace {
    bool on;
    const char *what;
    std::chrono::steady_clock::time_point t0, last;
    std::string line;
    HostTrace(const char *w, int debugLevel) : what(w)
    {
        on = debugLevel >= 2;
        if (on) t0 = last = std::chrono::steady_clock::now();
    }
    void mark(const char *name)
    {
        if (!on) return;
        auto t = std::chrono::steady_clock::now();
        char b[96];
        snprintf(b, sizeof b, " %s %.0f", name, std::chrono::duration<double, std::micro>(t - last).count());
        line += b;
        last = t;
    }
    ~HostTrace()
    {
        if (!on) return;
        fprintf(stderr, "[nda] %s host us:%s | total %.0f\n", what, line.c_str(),
                std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count());
    }
};

// ------------------------------------------------------------------------------------------
// several devices in ONE process: one host thread and one context per device, frames split
// [floor(g nF / G), floor((g+1) nF / G)), every device pulling only its frame columns over its own
// link; the partial histograms / float64 sums are merged by ONE ncclAllReduce on the compute streams
// (ncclCommInitAll; NCCL is dlopen()ed so the library carries no link-time dependency on it).
// within needs no collective: devices own disjoint frame columns of the mask.
// ------------------------------------------------------------------------------------------
struct NcclApi {
    bool tried = false, ok = false;
    void *handle = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::map<std::vector<int>, std::vector<ncclComm_t>> comms;
    std::mutex mu;
};
static NcclApi g_nccl;

static bool nccl_load()
{
    NcclApi &n = g_nccl;
    if (n.tried) return n.ok;
    n.tried = true;
    // the copy torch (or the application) has already mapped wins: two NCCL builds in one process do not mix
    for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
        n.handle = dlopen(name, RTLD_NOW | RTLD_NOLOAD);
        if (n.handle) break;
    }
    if (!n.handle)
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
    if (!n.handle) return false;
    n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(dlsym(n.handle, "ncclCommInitAll"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.handle, "ncclCommDestroy"));
    n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.handle, "ncclAllReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.handle, "ncclGetErrorString"));
    n.ok = n.CommInitAll && n.CommDestroy && n.AllReduce && n.GetErrorString;
    return n.ok;
}

// communicators of a device list, created on first use; nullptr: no NCCL (the caller sums on the host)
static const std::vector<ncclComm_t> *nccl_comms(const std::vector<int> &devs, const Config &cfg, int *rcOut)
{
    *rcOut = NDA_OK;
    if (cfg.collective == 2) return nullptr;
    std::lock_guard<std::mutex> lk(g_nccl.mu);
    if (!nccl_load()) {
        if (cfg.collective == 1) *rcOut = fail(NDA_ERR_CUDA, "NDA_COLLECTIVE=nccl but libnccl.so.2 could not be loaded: %s", dlerror());
        return nullptr;
    }
    auto it = g_nccl.comms.find(devs);
    if (it == g_nccl.comms.end()) {
        std::vector<ncclComm_t> cs(devs.size());
        const ncclResult_t r = g_nccl.CommInitAll(cs.data(), (int)devs.size(), devs.data());
        if (r != ncclSuccess) {
            if (cfg.collective == 1) *rcOut = fail(NDA_ERR_CUDA, "ncclCommInitAll failed: %s", g_nccl.GetErrorString(r));
            return nullptr;
        }
        it = g_nccl.comms.emplace(devs, std::move(cs)).first;
    }
    return &it->second;
}

struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, gen = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        const int g = gen;
        if (++waiting == n) { waiting = 0; ++gen; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};

// devices of the calling thread (nda_set_devices / nda_set_device / NDA_DEVICES / the current device)
static int devices_of_thread(std::vector<int> &out)
{
    if (g_devs.empty()) {
        static const std::vector<int> fromEnv = [] {
            std::vector<int> v;
            const char *e = getenv("NDA_DEVICES");
            if (!e || !*e) return v;
            int n = 0;
            if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return v; }
            if (!strcmp(e, "all")) { for (int i = 0; i < n; ++i) v.push_back(i); return v; }
            for (const char *p = e; *p;) {
                char *end = nullptr;
                const long d = strtol(p, &end, 10);
                if (end == p) break;
                if (d >= 0 && d < n && std::find(v.begin(), v.end(), (int)d) == v.end()) v.push_back((int)d);
                p = (*end == ',') ? end + 1 : end;
            }
            return v;
        }();
        if (!fromEnv.empty() && g_dev < 0) { g_devs = fromEnv; g_dev = g_devs[0]; }
        else {
            Ctx *c;
            const int rc = ctx_get(&c);
            if (rc) return rc;
            g_devs.assign(1, g_dev);
        }
    }
    out = g_devs;
    return NDA_OK;
}

struct Shard { int dev, fb, fe; };

// frames [fb, fe) over the thread's devices (no more devices than frames)
static int plan_shards(int fb, int fe, std::vector<Shard> &sh)
{
    std::vector<int> devs;
    const int rc = devices_of_thread(devs);
    if (rc) return rc;
    const int nFc = fe - fb;
    const int G = (int)std::max<size_t>(1, std::min<size_t>(devs.size(), (size_t)std::max(nFc, 1)));
    sh.clear();
    for (int g = 0; g < G; ++g)
        sh.push_back({devs[g], fb + (int)((long long)g * nFc / G), fb + (int)((long long)(g + 1) * nFc / G)});
    return NDA_OK;
}

// One PERSISTENT host thread per device (started on first use, parked on a condition variable between calls):
// a fresh std::thread per call would also make libgomp build a fresh thread team for the row gather on every
// call (its pools belong to the thread that opens the parallel region) -- measured: a 2-device call took 5.5 ms
// against 3.7 ms on one device before the workers were made persistent.
struct DeviceWorker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<void()> job;
    bool hasJob = false, done = false, quit = false, started = false;
    void loop()
    {
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return hasJob || quit; });
            if (quit) return;
            lk.unlock();
            job();
            lk.lock();
            hasJob = false;
            done = true;
            cv.notify_all();
        }
    }
    void post(std::function<void()> f)
    {
        std::unique_lock<std::mutex> lk(m);
        if (!started) { started = true; th = std::thread([this] { loop(); }); }
        job = std::move(f);
        done = false;
        hasJob = true;
        cv.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(m);
        cv.wait(lk, [&] { return done; });
    }
    void stop()
    {
        {
            std::unique_lock<std::mutex> lk(m);
            if (!started) return;
            quit = true;
            cv.notify_all();
        }
        th.join();
        started = false;
        quit = false;
    }
    ~DeviceWorker() { stop(); }
};
static DeviceWorker g_workers[16];
static std::mutex g_workers_mutex;                          // one multi-device call at a time uses the workers

// fn(g, ctx) runs on the worker of shard g's device, with that device current and its context locked; a
// single shard runs on the calling thread.  Returns the first failure (message copied into the caller's
// nda_last_error()).
template <class Fn>
static int run_shards(const std::vector<Shard> &sh, Fn fn)
{
    const int G = (int)sh.size();
    std::vector<int> rcs((size_t)G, NDA_OK);
    std::vector<std::string> errs((size_t)G);
    const int callerDev = g_dev;
    // every context exists before a worker starts: a device that cannot be set up fails the call here, not inside a
    // worker whose peers would then wait for it at the merge barrier
    for (int g = 0; g < G; ++g) {
        Ctx *c = nullptr;
        const int rc = ctx_for(sh[g].dev, &c);
        if (rc) { if (callerDev >= 0) cudaSetDevice(callerDev); return rc; }
    }
    auto body = [&](int g) {
        g_dev = sh[g].dev;
        Ctx *c = nullptr;
        int rc = ctx_for(sh[g].dev, &c);
        if (!rc) {
            std::lock_guard<std::mutex> lk(c->mu);
            begin_call(*c, G);
            rc = fn(g, *c);                                 // fn reaches the merge barrier on every path (merge_partials)
        }
        rcs[g] = rc;
        if (rc) errs[g] = g_err;
    };
    if (G == 1) body(0);
    else {
        std::lock_guard<std::mutex> lk(g_workers_mutex);
        for (int g = 0; g < G; ++g) g_workers[sh[g].dev].post([&body, g] { body(g); });
        for (int g = 0; g < G; ++g) g_workers[sh[g].dev].wait();
    }
    g_dev = callerDev;
    if (callerDev >= 0) cudaSetDevice(callerDev);
    for (int g = 0; g < G; ++g)
        if (rcs[g]) { g_err = errs[g]; return rcs[g]; }
    return NDA_OK;
}

// merge of per-device partials that sit in device buffers buf[g] (count elements of type T: u64 or f64) and
// delivery into `host`: ONE ncclAllReduce per device on its compute stream, then device 0 copies out; without
// NCCL every device copies its partial out and the host adds.  Called by every shard's thread.
struct Merge {
    const std::vector<ncclComm_t> *comms = nullptr;
    HostBarrier *bar = nullptr;
    std::atomic<int> failed{0};
    std::vector<std::vector<unsigned char>> parts;       // host-sum route
};

template <class T>
static int merge_partials(Merge &m, int g, int G, Ctx &c, cudaStream_t st, T *d_buf, size_t count, T *host, int rcSoFar)
{
    if (G == 1) {
        if (rcSoFar) return rcSoFar;
        CK(cudaMemcpyAsync(host, d_buf, count * sizeof(T), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        return NDA_OK;
    }
    if (rcSoFar) m.failed.store(1);
    m.bar->wait();                                         // every device's work is enqueued (or has failed)
    if (m.failed.load()) return rcSoFar ? rcSoFar : fail(NDA_ERR_CUDA, "another device of this call failed");
    if (m.comms) {
        NvtxRange nv("nda:allreduce", c.cfg.nvtx);
        const ncclDataType_t dt = sizeof(T) == 8 && std::is_floating_point<T>::value ? ncclFloat64 : ncclUint64;
        const ncclResult_t r = g_nccl.AllReduce(d_buf, d_buf, count, dt, ncclSum, (*m.comms)[g], st);
        if (r != ncclSuccess) return fail(NDA_ERR_CUDA, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r));
        if (g == 0) CK(cudaMemcpyAsync(host, d_buf, count * sizeof(T), cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        g_last_collective.store(1);
    } else {
        m.parts[g].resize(count * sizeof(T));
        cudaError_t e = cudaMemcpyAsync(m.parts[g].data(), d_buf, count * sizeof(T), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) m.failed.store(1);
        m.bar->wait();
        if (e != cudaSuccess) return fail(NDA_ERR_CUDA, "copy of a partial result failed: %s", cudaGetErrorString(e));
        if (m.failed.load()) return fail(NDA_ERR_CUDA, "another device of this call failed");
        if (g == 0) {
            for (size_t i = 0; i < count; ++i) {
                T acc = 0;
                for (int q = 0; q < G; ++q) acc += reinterpret_cast<const T *>(m.parts[q].data())[i];
                host[i] = acc;
            }
        }
        g_last_collective.store(2);
    }
    return NDA_OK;
}

// ------------------------------------------------------------------------------------------
// the three host paths on ONE device and ONE frame range (the multi-device entry points call them per shard)
// ------------------------------------------------------------------------------------------
// histogram of frames [fb, fe) accumulated into c.counts (device, zeroed here); stays on the device
static int rdf_shard(Ctx &c, const float *sel1, int n1, const float *sel2, int n2, const int *groupId, int nGroups,
                     int outSize, const float *cellDims, int nF, int fb, int fe, int sameSel, float dr, float *msOut)
{
    cudaStream_t st = c.stream;
    const size_t H = (size_t)nGroups * outSize;
    CK(c.counts.reserve(H * 8));
    CK(c.cell.reserve((size_t)nF * 12));
    CK(lane_acquire(c.lane[0], st));
    CK(lane_acquire(c.lane[1], c.lane[1].stream));
    CK(cudaMemsetAsync(c.counts.p, 0, H * 8, st));
    CK(reset_stats(c, st));
    if (fe <= fb) return NDA_OK;
    if (groupId) CK(c.gid.reserve((size_t)n2 * 4));
    {
        const SetupItem items[2] = {{c.cell.p, cellDims, (size_t)nF * 12},
                                    {groupId ? c.gid.p : nullptr, groupId, groupId ? (size_t)n2 * 4 : 0}};
        int rc = upload_setup(c, items, 2);
        if (rc) return rc;
        CK(cudaStreamWaitEvent(st, c.evSetupCopy, 0));
    }
    // frame chunks: the upload of chunk k+1 overlaps pack + kernel of chunk k
    Uploader up(c, nF, fb, fe);
    up.add(sel1, n1);
    if (!sameSel) up.add(sel2, n2);
    int rc = up.begin(8, c.cfg.uploadChunks);
    if (rc) return rc;
    CK(cudaEventRecord(c.ev0, st));
    CK(cudaEventRecord(c.evSetup, st));                    // lane 1 must see counts / cell / gid initialised
    CK(cudaStreamWaitEvent(c.lane[1].stream, c.evSetup, 0));
    rc = up.issue(0);
    if (rc) return rc;
    for (int k = 0; k < up.nChunks; ++k) {
        const int wc = up.f1(k) - up.f0(k);
        Ctx::Lane &L = c.lane[k & 1];                      // alternate lanes: kernel k+1 fills the tail of kernel k
        rc = up.acquire(k, L.stream);
        if (rc) return rc;
        rc = dev_rdf(c, L, L.stream, up.dev(0, k), n1, sameSel ? up.dev(0, k) : up.dev(1, k), n2,
                     groupId ? c.gid.as<int>() : nullptr, nGroups, c.counts.as<unsigned long long>(), outSize,
                     c.cell.as<float>() + (size_t)up.f0(k) * 3, wc, 0, wc, sameSel, dr, nullptr, nullptr,
                     cellDims + (size_t)up.f0(k) * 3);
        if (rc) return rc;
        rc = up.release(k, L.stream);
        if (rc) return rc;
        if (k + 1 < up.nChunks) { rc = up.issue(k + 1); if (rc) return rc; }
    }
    CK(lane_release(c.lane[1], c.lane[1].stream));
    CK(cudaEventRecord(c.evLaneJoin, c.lane[1].stream));
    CK(cudaStreamWaitEvent(st, c.evLaneJoin, 0));
    CK(lane_release(c.lane[0], st));
    CK(cudaEventRecord(c.ev1, st));
    if (msOut) {
        CK(cudaEventSynchronize(c.ev1));
        CK(cudaEventElapsedTime(msOut, c.ev0, c.ev1));
    }
    return NDA_OK;
}

// float64 sums of frames [fb, fe) accumulated into c.sum (device, zeroed here); stays on the device
static int distances_shard(Ctx &c, const float *sel1, int n1, const float *sel2, int n2, const float *cellDims,
                           int nF, int fb, int fe, int sameSel)
{
    cudaStream_t st = c.stream;
    const size_t M = (size_t)n1 * n2;
    CK(c.sum.reserve(M * 8));
    CK(c.cell.reserve((size_t)nF * 12));
    CK(cudaMemsetAsync(c.sum.p, 0, M * 8, st));
    if (fe <= fb) return NDA_OK;
    {
        const SetupItem items[1] = {{c.cell.p, cellDims, (size_t)nF * 12}};
        int rc = upload_setup(c, items, 1);
        if (rc) return rc;
        CK(cudaStreamWaitEvent(st, c.evSetupCopy, 0));
    }
    Uploader up(c, nF, fb, fe);
    up.add(sel1, n1);
    if (!sameSel) up.add(sel2, n2);
    int rc = up.begin(8, c.cfg.uploadChunks);
    if (rc) return rc;
    rc = up.issue(0);
    if (rc) return rc;
    for (int k = 0; k < up.nChunks; ++k) {
        const int wc = up.f1(k) - up.f0(k);
        rc = up.acquire(k, st);
        if (rc) return rc;
        NvtxRange nv("nda:distances_kernel", c.cfg.nvtx);
        rc = dev_distances(c, st, up.dev(0, k), n1, sameSel ? up.dev(0, k) : up.dev(1, k), n2,
                           c.sum.as<double>(), c.cell.as<float>() + (size_t)up.f0(k) * 3, wc, 0, wc, sameSel);
        if (rc) return rc;
        rc = up.release(k, st);
        if (rc) return rc;
        if (k + 1 < up.nChunks) { rc = up.issue(k + 1); if (rc) return rc; }
    }
    return NDA_OK;
}

// cut-off bitmask of frames [fb, fe): rows [0, fe - fb) of `bits` (may be NULL) and columns [fb, fe) of `out`
// (may be NULL) are written; everything is back on the host when this returns
static int within_shard(Ctx &c, const float *allAtoms, int nAtoms, int nF, const int *refSel, int refSize,
                        const int *outSel, int outSelSize, uint32_t *bits, int *out, const float *cellDims,
                        float distance, int fb, int fe)
{
    cudaStream_t st = c.stream;
    HostTrace ht("within", c.cfg.debug);
    const int nFc = fe - fb;
    const int words = (outSelSize + 31) / 32;
    if (nFc <= 0) return NDA_OK;

    // ---- which rows travel, and the selections re-expressed in uploaded-row numbers
    Ctx::WithinPlan &pl = c.plan;
    const bool samePlan = pl.nAtoms == nAtoms && (int)pl.refSel.size() == refSize && (int)pl.outSel.size() == outSelSize &&
                          (refSize == 0 || memcmp(pl.refSel.data(), refSel, (size_t)refSize * 4) == 0) &&
                          memcmp(pl.outSel.data(), outSel, (size_t)outSelSize * 4) == 0;
    if (!samePlan) {
        pl.nAtoms = -1;
        plan_rows(refSel, refSize, outSel, outSelSize, nAtoms, pl.remap, pl.runs, pl.nRows);
        pl.ref2.resize((size_t)std::max(refSize, 1));
        pl.out2.resize((size_t)outSelSize);
        for (int i = 0; i < refSize; ++i) pl.ref2[i] = pl.remap[refSel[i]];
        for (int i = 0; i < outSelSize; ++i) pl.out2[i] = pl.remap[outSel[i]];
        pl.refSel.assign(refSel, refSel + refSize);
        pl.outSel.assign(outSel, outSel + outSelSize);
        pl.nAtoms = nAtoms;
    }
    const std::vector<int> &ref2 = pl.ref2, &out2 = pl.out2;
    const int nRows = pl.nRows;

    // ---- frame chunks: copy of chunk k+1 overlaps pack + kernel of chunk k (two raw buffers)
    Uploader up(c, nF, fb, fe);
    up.add(allAtoms, nRows, pl.runs);
    int rc = up.begin(6, c.cfg.withinChunks);
    if (rc) return rc;
    const int nChunks = up.nChunks;
    ht.mark("plan");
    CK(c.cell.reserve((size_t)nF * 12));
    CK(c.idx1.reserve((size_t)std::max(refSize, 1) * 4));
    CK(c.idx2.reserve((size_t)outSelSize * 4));
    {                                                       // ~60 KB on the copy stream, then the first chunk
        const SetupItem items[3] = {{c.cell.p, cellDims, (size_t)nF * 12},
                                    {c.idx1.p, ref2.data(), (size_t)refSize * 4},
                                    {c.idx2.p, out2.data(), (size_t)outSelSize * 4}};
        rc = upload_setup(c, items, 3);
        if (rc) return rc;
    }
    rc = up.issue(0);                                       // the link is the bottleneck: start it first
    if (rc) return rc;
    ht.mark("issue0");

    CK(c.bits.reserve((size_t)nFc * words * 4));
    CK(c.pinned_reserve((size_t)nFc * words * 4));
    uint32_t *hbits = reinterpret_cast<uint32_t *>(c.pinned);

    CK(lane_acquire(c.lane[0], st));
    CK(lane_acquire(c.lane[1], c.lane[1].stream));
    CK(reset_stats(c, st));
    CK(cudaStreamWaitEvent(st, c.evSetupCopy, 0));
    CK(cudaEventRecord(c.ev0, st));

    // host: set the 1s of frames [f0, f1) (relative to fb).  ~6e5 cache-missing stores for config 2, spread over threads:
    // part p of `parts` takes every parts-th frame block
    auto scatter_part = [&](int f0, int f1, int part, int parts) {
        if (!out) return;
        const int n = f1 - f0;
        const int a = f0 + (int)((long long)n * part / parts), b = f0 + (int)((long long)n * (part + 1) / parts);
        for (int f = a; f < b; ++f) {
            const uint32_t *row = hbits + (size_t)f * words;
            for (int w = 0; w < words; ++w) {
                uint32_t v = row[w];
                while (v) {                                  // the reference only ever stores 1 (getWithin.cpp:77,83)
                    const int bpos = __builtin_ctz(v);
                    v &= v - 1;
                    out[(size_t)outSel[w * 32 + bpos] * nF + fb + f] = 1;
                }
            }
        }
    };
    const int hostThreads = c.hostThreads;
    (void)hostThreads;                                      // (used by the OpenMP pragma below)
    auto scatter = [&](int k) {                             // chunk k with the whole team
        if (!out) return;
        const int f0 = up.f0(k) - fb, f1 = up.f1(k) - fb;
        #pragma omp parallel num_threads(hostThreads) if (f1 - f0 >= 16)
        scatter_part(f0, f1, omp_get_thread_num(), omp_get_num_threads());
    };

    CK(cudaEventRecord(c.evSetup, st));                     // lane 1 must see cell / selections uploaded
    CK(cudaStreamWaitEvent(c.lane[1].stream, c.evSetup, 0));
    ht.mark("setup");
    for (int k = 0; k < nChunks; ++k) {
        const int f0 = up.f0(k), wc = up.f1(k) - f0, buf = k & 1;
        Ctx::Lane &L = c.lane[buf];                         // alternate lanes: kernel k+1 fills the tail of kernel k
        cudaStream_t ls = L.stream;
        rc = up.acquire(k, ls);
        if (rc) return rc;
        rc = dev_within(c, L, ls, up.dev(0, k), nRows, wc, c.idx1.as<int>(), refSize, c.idx2.as<int>(), outSelSize,
                        c.bits.as<unsigned>() + (size_t)(f0 - fb) * words, nullptr,
                        c.cell.as<float>() + (size_t)f0 * 3, distance, 0, wc, cellDims + (size_t)f0 * 3);
        if (rc) return rc;
        rc = up.release(k, ls);
        if (rc) return rc;
        if (c.cfg.d2hCopyEngine) {                           // A/B switch: the old way
            CK(cudaMemcpyAsync(hbits + (size_t)(f0 - fb) * words, c.bits.as<unsigned>() + (size_t)(f0 - fb) * words,
                               (size_t)wc * words * 4, cudaMemcpyDeviceToHost, ls));
        } else {                                             // SMs write the pinned buffer: no copy-engine queueing
            const size_t nW = (size_t)wc * words;
            copy_words_kernel<<<(unsigned)std::min<size_t>((nW + 255) / 256, 296), 256, 0, ls>>>(
                c.bits.as<unsigned>() + (size_t)(f0 - fb) * words, hbits + (size_t)(f0 - fb) * words, nW);
            LAUNCHED();
        }
        CK(cudaEventRecord(c.evBits[buf], ls));
        // host: gather chunk k+1 while the GPU works on k, and -- beside the gather, on a quarter of the team -- set the 1s
        // of chunk k-1, whose bitmask has long arrived
        if (k > 0) {
            CK(cudaEventSynchronize(c.evBits[buf ^ 1]));
            ht.mark("wait");
        }
        if (k + 1 < nChunks) {
            if (k > 0 && out) {
                const int sf0 = up.f0(k - 1) - fb, sf1 = up.f1(k - 1) - fb;
                const std::function<void(int, int)> side = [&, sf0, sf1](int part, int parts) { scatter_part(sf0, sf1, part, parts); };
                rc = up.issue(k + 1, &side);
            } else {
                rc = up.issue(k + 1);
            }
            if (rc) return rc;
            ht.mark("gather+scat");
        } else if (k > 0) {
            scatter(k - 1);
            ht.mark("scat");
        }
    }
    CK(lane_release(c.lane[1], c.lane[1].stream));
    CK(cudaEventRecord(c.evLaneJoin, c.lane[1].stream));
    CK(cudaStreamWaitEvent(st, c.evLaneJoin, 0));
    CK(lane_release(c.lane[0], st));
    CK(cudaEventRecord(c.ev1, st));
    ht.mark("loop");
    CK(cudaStreamSynchronize(st));
    ht.mark("sync");
    scatter(nChunks - 1);
    ht.mark("scatter");
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    rc = read_stats(c, (double)refSize * outSelSize * nFc, ms);
    if (rc) return rc;
    if (bits) memcpy(bits, hbits, (size_t)nFc * words * 4);
    return NDA_OK;
}

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
// lock the calling thread's (first) device for a device-pointer / single-device call
#define NDA_ENTER()                                                                           \
    Ctx *c;                                                                                   \
    int rc = ctx_get(&c);                                                                     \
    if (rc) return rc;                                                                        \
    std::lock_guard<std::mutex> lk(c->mu);                                                    \
    begin_call(*c)

extern "C" {

const char *nda_last_error(void) { return g_err.c_str(); }
int nda_get_parallel_backend(void) { return 2; }

int nda_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int nda_set_device(int device)
{
    int n = nda_device_count();
    if (device < 0 || device >= n) return fail(NDA_ERR_ARG, "device %d out of range (%d visible)", device, n);
    g_dev = device;
    g_devs.assign(1, device);
    CK(cudaSetDevice(device));
    return NDA_OK;
}

int nda_set_devices(const int *devices, int n)
{
    const int nd = nda_device_count();
    if (!devices || n < 1 || n > 16) return fail(NDA_ERR_ARG, "need 1..16 devices");
    std::vector<int> v;
    for (int i = 0; i < n; ++i) {
        if (devices[i] < 0 || devices[i] >= nd) return fail(NDA_ERR_ARG, "device %d out of range (%d visible)", devices[i], nd);
        if (std::find(v.begin(), v.end(), devices[i]) != v.end()) return fail(NDA_ERR_ARG, "device %d listed twice", devices[i]);
        v.push_back(devices[i]);
    }
    g_devs = v;
    g_dev = v[0];
    CK(cudaSetDevice(g_dev));
    return NDA_OK;
}

int nda_get_devices(int *devices, int capacity)
{
    std::vector<int> v;
    const int rc = devices_of_thread(v);
    if (rc) return -1;
    for (int i = 0; i < (int)v.size() && i < capacity; ++i) devices[i] = v[i];
    return (int)v.size();
}

int nda_get_device(void) { return g_dev; }
int nda_last_filter(void) { return g_last_filter.load(); }
int nda_last_collective(void) { return g_last_collective.load(); }

int nda_set_option(const char *name, const char *value)
{
    if (!name) return fail(NDA_ERR_ARG, "null option name");
    config_init();
    std::lock_guard<std::mutex> lk(g_cfg_mutex);
    if (!apply_option(g_cfg, name, value)) return fail(NDA_ERR_ARG, "unknown option %s", name);
    return NDA_OK;
}

// releases every device buffer, pinned staging area, stream, event and communicator the library holds;
// the next call builds them again
int nda_shutdown(void)
{
    {
        std::lock_guard<std::mutex> lk(g_workers_mutex);
        for (auto &w : g_workers) w.stop();
    }
    {
        std::lock_guard<std::mutex> lk(g_nccl.mu);
        for (auto &kv : g_nccl.comms)
            for (ncclComm_t cm : kv.second) g_nccl.CommDestroy(cm);
        g_nccl.comms.clear();
    }
    std::lock_guard<std::mutex> lkInit(g_ctx_init_mutex);
    for (int d = 0; d < 16; ++d) {
        Ctx &c = g_ctx[d];
        if (!c.ready) continue;
        std::lock_guard<std::mutex> lk(c.mu);
        if (cudaSetDevice(d) != cudaSuccess) { cudaGetLastError(); continue; }
        cudaDeviceSynchronize();
        auto freeBuf = [](DevBuf &b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; };
        for (auto &L : c.lane) {
            for (DevBuf *b : {&L.packA, &L.packB, &L.packAw, &L.packBw, &L.fp, &L.maxAbs, &L.ctl, &L.embA, &L.embB, &L.thrDot}) freeBuf(*b);
            if (L.evBusy) cudaEventDestroy(L.evBusy);
            L.evBusy = nullptr;
        }
        for (auto &r : c.raw) for (auto &b : r) freeBuf(b);
        for (DevBuf *b : {&c.cell, &c.idx1, &c.idx2, &c.gid, &c.counts, &c.bits, &c.sum}) freeBuf(*b);
        for (void **h : {&c.pinned, &c.setupStage, &c.stage[0], &c.stage[1]}) { if (*h) cudaFreeHost(*h); *h = nullptr; }
        c.pinnedCap = c.setupCap = c.stageCap[0] = c.stageCap[1] = 0;
        for (cudaEvent_t *e : {&c.evJoin, &c.ev0, &c.ev1, &c.evCopied[0], &c.evCopied[1], &c.evDone[0], &c.evDone[1],
                               &c.evBits[0], &c.evBits[1], &c.evSetup, &c.evLaneJoin, &c.evSetupCopy}) {
            if (*e) cudaEventDestroy(*e);
            *e = nullptr;
        }
        for (cudaEvent_t e : c.kev) cudaEventDestroy(e);
        c.kev.clear();
        c.kevUsed = 0;
        if (c.lane[1].stream) cudaStreamDestroy(c.lane[1].stream);
        for (cudaStream_t *s : {&c.stream, &c.copyStream, &c.copyStream2}) { if (*s) cudaStreamDestroy(*s); *s = nullptr; }
        c.lane[0].stream = c.lane[1].stream = nullptr;
        c.plan = Ctx::WithinPlan();
        c.hostCell.clear();
        c.hostCellPtr = nullptr;
        c.statsPending = false;
        c.ready = false;
    }
    return NDA_OK;
}

// debugging aid (trace builds, tools/): copies the clock64 trace of the last within call, 256 x 32 values
int nda_debug_tc_trace(long long *out)
{
    NDA_ENTER();
    if (!c->sum.p || c->sum.cap < 256 * 32 * 8) return fail(NDA_ERR_ARG, "no trace recorded");
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, c->sum.p, 256 * 32 * 8, cudaMemcpyDeviceToHost));
    return NDA_OK;
}
uint64_t nda_launch_count(void) { return g_launches.load(); }
void nda_launch_count_reset(void) { g_launches.store(0); }
int nda_set_within_pure_predicate(int on) { g_within_pure.store(on ? 1 : 0); return NDA_OK; }

int nda_last_stats(double stats[4])
{
    Ctx *c;
    int rc = ctx_get(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    rc = fetch_stats(*c);
    if (rc) return rc;
    memcpy(stats, c->stats, sizeof c->stats);
    return NDA_OK;
}

int nda_last_kernel_ms(double *ms, int *launches)
{
    if (!ms) return fail(NDA_ERR_ARG, "null pointer argument");
    Ctx *c;
    int rc = ctx_get(&c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    double tot = 0.0;
    int n = 0;
    for (int i = 0; i + 1 < c->kevUsed; i += 2) {
        float t = 0.f;
        CK(cudaEventSynchronize(c->kev[i + 1]));
        CK(cudaEventElapsedTime(&t, c->kev[i], c->kev[i + 1]));
        tot += t;
        ++n;
    }
    *ms = tot;
    if (launches) *launches = n;
    return NDA_OK;
}

// ---- device-pointer entry points -----------------------------------------------------------
// All of them share the context's lane-0 scratch buffers; a per-lane event orders consecutive calls even
// when they arrive on different caller streams.
int nda_dev_radial_nbr_counts(const float *d_sel1, int n1, const float *d_sel2, int n2,
                              const int *d_groupId, int nGroups, unsigned long long *d_counts, int outSize,
                              const float *d_cellDims, int nF, int fb, int fe, int sameSel, float dr, void *stream)
{
    if (!d_sel1 || !d_sel2 || !d_counts || !d_cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (n1 < 0 || n2 < 0 || outSize < 0 || nF < 0 || fb < 0 || fe > nF) return fail(NDA_ERR_ARG, "bad size / frame range");
    if (sameSel && (d_sel1 != d_sel2 || n1 != n2)) return fail(NDA_ERR_ARG, "sameSel requires sel2 == sel1");
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CK(lane_acquire(c->lane[0], st));
    CK(reset_stats(*c, st));
    rc = dev_rdf(*c, c->lane[0], st, d_sel1, n1, d_sel2, n2, d_groupId, nGroups, d_counts, outSize, d_cellDims, nF, fb, fe, sameSel, dr);
    if (rc) return rc;
    CK(lane_release(c->lane[0], st));
    return read_stats(*c, (sameSel ? 0.5 * (double)n1 * ((double)n1 - 1.0) : (double)n1 * (double)n2) * (fe - fb), 0.f);
}

int nda_dev_within(const float *d_allAtoms, int nAtoms, int nF, const int *d_refSel, int refSize,
                   const int *d_outSel, int outSelSize, uint32_t *d_bits, int *d_out,
                   const float *d_cellDims, float distance, int fb, int fe, void *stream)
{
    if (!d_allAtoms || !d_bits || !d_cellDims || (refSize > 0 && !d_refSel) || (outSelSize > 0 && !d_outSel))
        return fail(NDA_ERR_ARG, "null pointer argument");
    if (nAtoms < 0 || nF < 0 || refSize < 0 || outSelSize < 0 || fb < 0 || fe > nF) return fail(NDA_ERR_ARG, "bad size / frame range");
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CK(lane_acquire(c->lane[0], st));
    CK(reset_stats(*c, st));
    rc = dev_within(*c, c->lane[0], st, d_allAtoms, nAtoms, nF, d_refSel, refSize, d_outSel, outSelSize, d_bits, d_out,
                    d_cellDims, distance, fb, fe);
    if (rc) return rc;
    CK(lane_release(c->lane[0], st));
    return read_stats(*c, (double)refSize * outSelSize * (fe - fb), 0.f);
}

int nda_dev_distances_sum(const float *d_sel1, int n1, const float *d_sel2, int n2, double *d_sum,
                          const float *d_cellDims, int nF, int fb, int fe, int sameSel, void *stream)
{
    if (!d_sel1 || !d_sel2 || !d_sum || !d_cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (n1 < 0 || n2 < 0 || nF < 0 || fb < 0 || fe > nF) return fail(NDA_ERR_ARG, "bad size / frame range");
    if (sameSel && n1 != n2) return fail(NDA_ERR_ARG, "sameSel requires sel2 == sel1");
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return dev_distances(*c, st, d_sel1, n1, d_sel2, n2, d_sum, d_cellDims, nF, fb, fe, sameSel);
}

// Selections given as row indices into ONE resident all-atom array (a trajectory uploaded once and
// queried many times: ResidueWiseWaterDensity residue by residue, ResidenceTime per time origin, ...)
int nda_dev_radial_nbr_counts_sel(const float *d_all, int nAtoms, const int *d_idx1, int n1, const int *d_idx2, int n2,
                                  const int *d_groupId, int nGroups, unsigned long long *d_counts, int outSize,
                                  const float *d_cellDims, int nF, int fb, int fe, int sameSel, float dr, void *stream)
{
    if (!d_all || !d_idx1 || !d_idx2 || !d_counts || !d_cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (nAtoms < 0 || n1 < 0 || n2 < 0 || outSize < 0 || nF < 0 || fb < 0 || fe > nF) return fail(NDA_ERR_ARG, "bad size / frame range");
    if (sameSel && (d_idx1 != d_idx2 || n1 != n2)) return fail(NDA_ERR_ARG, "sameSel requires idx2 == idx1");
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    CK(lane_acquire(c->lane[0], st));
    CK(reset_stats(*c, st));
    rc = dev_rdf(*c, c->lane[0], st, d_all, n1, d_all, n2, d_groupId, nGroups, d_counts, outSize, d_cellDims, nF, fb, fe, sameSel, dr,
                 d_idx1, d_idx2);
    if (rc) return rc;
    CK(lane_release(c->lane[0], st));
    return read_stats(*c, (sameSel ? 0.5 * (double)n1 * ((double)n1 - 1.0) : (double)n1 * (double)n2) * (fe - fb), 0.f);
}

int nda_dev_distances_sum_sel(const float *d_all, int nAtoms, const int *d_idx1, int n1, const int *d_idx2, int n2,
                              double *d_sum, const float *d_cellDims, int nF, int fb, int fe, int sameSel, void *stream)
{
    if (!d_all || !d_idx1 || !d_idx2 || !d_sum || !d_cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (nAtoms < 0 || n1 < 0 || n2 < 0 || nF < 0 || fb < 0 || fe > nF) return fail(NDA_ERR_ARG, "bad size / frame range");
    if (sameSel && n1 != n2) return fail(NDA_ERR_ARG, "sameSel requires idx2 == idx1");
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    return dev_distances(*c, st, d_all, n1, d_all, n2, d_sum, d_cellDims, nF, fb, fe, sameSel, d_idx1, d_idx2);
}

// ---- host entry points ---------------------------------------------------------------------
int nda_get_radial_nbr_counts(const float *sel1, int n1, const float *sel2, int n2,
                              const int *groupId, int nGroups, uint64_t *counts, float *out, int outSize,
                              const float *cellDims, int nF, int fb, int fe, int sameSel, float dr)
{
    if (!sel1 || !sel2 || !cellDims || (!counts && !out)) return fail(NDA_ERR_ARG, "null pointer argument");
    if (n1 < 0 || n2 < 0 || outSize < 0 || nF <= 0 || fb < 0 || fe > nF || fb > fe) return fail(NDA_ERR_ARG, "bad size / frame range");
    if (!groupId) nGroups = 1;
    if (nGroups < 1) return fail(NDA_ERR_ARG, "nGroups must be >= 1");
    if (sameSel && n1 != n2) return fail(NDA_ERR_ARG, "sameSel requires sel2 to be sel1");
    if (groupId)
        for (int i = 0; i < n2; ++i)
            if (groupId[i] < 0 || groupId[i] >= nGroups) return fail(NDA_ERR_ARG, "groupId[%d] = %d outside [0, %d)", i, groupId[i], nGroups);
    const size_t H = (size_t)nGroups * outSize;
    std::vector<unsigned long long> h(H, 0ull);
    if (H && n1 && n2 && fe > fb) {
        std::vector<Shard> sh;
        int rc = plan_shards(fb, fe, sh);
        if (rc) return rc;
        const int G = (int)sh.size();
        Merge m;
        HostBarrier bar(G);
        m.bar = &bar;
        m.parts.resize((size_t)G);
        g_last_collective.store(0);
        if (G > 1) {
            std::vector<int> devs;
            for (const Shard &s : sh) devs.push_back(s.dev);
            m.comms = nccl_comms(devs, config_get(), &rc);
            if (rc) return rc;
        }
        const double perFrame = sameSel ? 0.5 * (double)n1 * ((double)n1 - 1.0) : (double)n1 * (double)n2;
        rc = run_shards(sh, [&](int g, Ctx &c) -> int {
            float ms = 0.f;
            int r = rdf_shard(c, sel1, n1, sel2, n2, groupId, nGroups, outSize, cellDims, nF, sh[g].fb, sh[g].fe, sameSel, dr,
                              G == 1 ? nullptr : &ms);
            r = merge_partials<unsigned long long>(m, g, G, c, c.stream, c.counts.as<unsigned long long>(), H, h.data(), r);
            if (r) return r;
            if (G == 1) CK(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
            return read_stats(c, perFrame * (sh[g].fe - sh[g].fb), ms);
        });
        if (rc) return rc;
    }
    if (counts) for (size_t i = 0; i < H; ++i) counts[i] = (uint64_t)h[i];
    if (out)    for (size_t i = 0; i < H; ++i) out[i] = (float)((double)h[i] / (double)nF);
    return NDA_OK;
}

int nda_get_radial_nbr_density(const float *sel1, int n1, const float *sel2, int n2, float *out, int outSize,
                               const float *cellDims, int nF, int sameSel, float maxR, float dr)
{
    (void)maxR;                                             // unused by the reference as well (SURVEY B-9)
    if (!out) return fail(NDA_ERR_ARG, "null pointer argument");
    if (outSize < 0) return fail(NDA_ERR_ARG, "bad size");
    // the reference ACCUMULATES into out (getRadialNbrDensity.cpp:40-41: out[idx] += 1 / nbrFrames per pair);
    // here the integer count is divided once and added once -- equal for the zero-initialised array every stock
    // caller passes (RadialDensity.py:205, 76), and a caller that reuses `out` still gets a sum
    std::vector<float> d((size_t)outSize, 0.f);
    const int rc = nda_get_radial_nbr_counts(sel1, n1, sel2, n2, nullptr, 1, nullptr, d.data(), outSize, cellDims, nF, 0, nF, sameSel, dr);
    if (rc) return rc;
    for (int i = 0; i < outSize; ++i) out[i] += d[i];
    return NDA_OK;
}

int nda_get_distances_sum(const float *sel1, int n1, const float *sel2, int n2, double *sum,
                          const float *cellDims, int nF, int fb, int fe, int sameSel)
{
    if (!sel1 || !sel2 || !sum || !cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (n1 < 0 || n2 < 0 || nF <= 0 || fb < 0 || fe > nF || fb > fe) return fail(NDA_ERR_ARG, "bad size / frame range");
    if (sameSel && n1 != n2) return fail(NDA_ERR_ARG, "sameSel requires sel2 to be sel1");
    const size_t M = (size_t)n1 * n2;
    if (M == 0) return NDA_OK;
    if (fe == fb) { memset(sum, 0, M * 8); return NDA_OK; }
    std::vector<Shard> sh;
    int rc = plan_shards(fb, fe, sh);
    if (rc) return rc;
    const int G = (int)sh.size();
    Merge m;
    HostBarrier bar(G);
    m.bar = &bar;
    m.parts.resize((size_t)G);
    g_last_collective.store(0);
    if (G > 1) {
        std::vector<int> devs;
        for (const Shard &s : sh) devs.push_back(s.dev);
        m.comms = nccl_comms(devs, config_get(), &rc);
        if (rc) return rc;
    }
    return run_shards(sh, [&](int g, Ctx &c) -> int {
        int r = distances_shard(c, sel1, n1, sel2, n2, cellDims, nF, sh[g].fb, sh[g].fe, sameSel);
        if (!r && sameSel && G == 1) {
            dim3 grid((n1 + 255) / 256, std::min(n1, 4096));
            symmetrize_kernel<<<grid, 256, 0, c.stream>>>(c.sum.as<double>(), n1);
            g_launches.fetch_add(1, std::memory_order_relaxed);
            const cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) r = fail(NDA_ERR_CUDA, "symmetrize_kernel launch failed: %s", cudaGetErrorString(e));
        }
        r = merge_partials<double>(m, g, G, c, c.stream, c.sum.as<double>(), M, sum, r);
        if (r) return r;
        if (sameSel && G > 1 && g == 0)                     // mirror the merged upper triangle (cheap, host side)
            for (int i = 0; i < n1; ++i)
                for (int j = i + 1; j < n1; ++j) sum[(size_t)j * n1 + i] = sum[(size_t)i * n1 + j];
        return NDA_OK;
    });
}

int nda_get_distances(const float *sel1, int n1, const float *sel2, int n2, float *out,
                      const float *cellDims, int nF, int sameSel)
{
    if (!out) return fail(NDA_ERR_ARG, "null pointer argument");
    if (n1 < 0 || n2 < 0) return fail(NDA_ERR_ARG, "bad size");
    std::vector<double> s((size_t)n1 * n2, 0.0);
    int rc = nda_get_distances_sum(sel1, n1, sel2, n2, s.data(), cellDims, nF, 0, nF, sameSel);
    if (rc) return rc;
    for (size_t i = 0; i < s.size(); ++i) out[i] = (float)(s[i] / (double)nF);
    return NDA_OK;
}

int nda_get_within_bits(const float *allAtoms, int nAtoms, int nF, const int *refSel, int refSize,
                        const int *outSel, int outSelSize, uint32_t *bits, int *out,
                        const float *cellDims, float distance, int fb, int fe)
{
    if (!allAtoms || !cellDims || (!bits && !out) || (refSize > 0 && !refSel) || (outSelSize > 0 && !outSel))
        return fail(NDA_ERR_ARG, "null pointer argument");
    if (nAtoms < 0 || nF <= 0 || refSize < 0 || outSelSize < 0 || fb < 0 || fe > nF || fb > fe)
        return fail(NDA_ERR_ARG, "bad size / frame range");
    for (int i = 0; i < refSize; ++i)
        if (refSel[i] < 0 || refSel[i] >= nAtoms) return fail(NDA_ERR_ARG, "refSel[%d] = %d outside [0, %d)", i, refSel[i], nAtoms);
    for (int i = 0; i < outSelSize; ++i)
        if (outSel[i] < 0 || outSel[i] >= nAtoms) return fail(NDA_ERR_ARG, "outSel[%d] = %d outside [0, %d)", i, outSel[i], nAtoms);
    const Config cfg = config_get();
    // SURVEY B-1: the legacy CUDA build (lib/cuda/src/getWithin.cu:27-28) marks every refSel atom in every
    // frame whether or not it belongs to outSel; the OpenMP build (the oracle) does not.  Opt-in.
    if (cfg.compatCuda && out && fe > fb) {
        for (int i = 0; i < refSize; ++i)
            for (int f = fb; f < fe; ++f) out[(size_t)refSel[i] * nF + f] = 1;
    }
    if (outSelSize == 0 || fe == fb) return NDA_OK;
    const int words = (outSelSize + 31) / 32;
    std::vector<Shard> sh;
    int rc = plan_shards(fb, fe, sh);
    if (rc) return rc;
    g_last_collective.store(0);
    return run_shards(sh, [&](int g, Ctx &c) -> int {
        return within_shard(c, allAtoms, nAtoms, nF, refSel, refSize, outSel, outSelSize,
                            bits ? bits + (size_t)(sh[g].fb - fb) * words : nullptr, out, cellDims, distance, sh[g].fb, sh[g].fe);
    });
}

int nda_get_within(const float *allAtoms, int nAtoms, int nF, const int *refSel, int refSize,
                   const int *outSel, int outSelSize, int *out, const float *cellDims, float distance)
{
    if (!out) return fail(NDA_ERR_ARG, "null pointer argument");
    return nda_get_within_bits(allAtoms, nAtoms, nF, refSel, refSize, outSel, outSelSize, nullptr, out,
                               cellDims, distance, 0, nF);
}

// ---- water at the protein surface ----------------------------------------------
// Both calls work IN PLACE on a host array, like the reference: upload, pack the protein frames into
// 16-byte tiles, one kernel per pass of frames, download.
static int surface_common(Ctx &c, cudaStream_t st, const float *prot, int sizeP, const float *cellDims, int nF,
                          float **d_prot, float **d_cell)
{
    CK(c.raw[0][1].reserve((size_t)sizeP * nF * 12 + 16));
    CK(c.cell.reserve((size_t)nF * 12));
    CK(cudaMemcpyAsync(c.raw[0][1].p, prot, (size_t)sizeP * nF * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c.cell.p, cellDims, (size_t)nF * 12, cudaMemcpyHostToDevice, st));
    *d_prot = c.raw[0][1].as<float>();
    *d_cell = c.cell.as<float>();
    return NDA_OK;
}

int nda_set_water_dist_pbc(float *water, int sizeW, const float *prot, int sizeP,
                           const float *cellDims, int nF, int nbrWAtoms)
{
    if (!water || !prot || !cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (sizeW < 0 || sizeP < 0 || nF <= 0 || nbrWAtoms <= 0) return fail(NDA_ERR_ARG, "bad size");
    if (sizeW == 0) return NDA_OK;
    NDA_ENTER();
    cudaStream_t st = c->stream;
    Ctx::Lane &L = c->lane[0];
    const size_t wBytes = (size_t)sizeW * nbrWAtoms * nF * 12;
    CK(c->raw[0][0].reserve(wBytes + 16));
    CK(cudaMemcpyAsync(c->raw[0][0].p, water, wBytes, cudaMemcpyHostToDevice, st));
    float *d_prot = nullptr, *d_cell = nullptr;
    rc = surface_common(*c, st, prot, sizeP, cellDims, nF, &d_prot, &d_cell);
    if (rc) return rc;
    const int nPPad = round_up(std::max(sizeP, 1), 32);
    const int pass = std::min(32768, frames_per_pass((size_t)nPPad * sizeof(float4), nF));
    CK(L.packB.reserve((size_t)nPPad * pass * sizeof(float4)));
    CK(L.maxAbs.reserve((size_t)pass * sizeof(unsigned)));
    for (int f0 = 0; f0 < nF; f0 += pass) {
        const int nFc = std::min(pass, nF - f0);
        CK(cudaMemsetAsync(L.maxAbs.p, 0, (size_t)nFc * sizeof(unsigned), st));
        rc = launch_pack(*c, st, d_prot, nF, nullptr, sizeP, nPPad, f0, nFc, nullptr, L.packB.as<float4>(),
                         L.maxAbs.as<unsigned>(), d_cell, nullptr);
        if (rc) return rc;
        dim3 grid((sizeW + SURF_THREADS - 1) / SURF_THREADS, nFc);
        CK(c->kev_record(st));
        water_nearest_shift_kernel<<<grid, SURF_THREADS, 0, st>>>(c->raw[0][0].as<float>(), sizeW, nbrWAtoms, nF, f0,
                                                                  L.packB.as<float4>(), sizeP, nPPad, d_cell);
        LAUNCHED();
        CK(c->kev_record(st));
    }
    CK(cudaMemcpyAsync(water, c->raw[0][0].p, wBytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return NDA_OK;
}

int nda_water_orient_at_surface(float *waterO, int sizeO, const float *watVec, const float *prot, int sizeP,
                                float *out_unused, const float *cellDims, int nF, float minR, float maxR, int maxN)
{
    (void)out_unused;                                       // unused by the reference as well (libFunc.h)
    if (!waterO || !watVec || !prot || !cellDims) return fail(NDA_ERR_ARG, "null pointer argument");
    if (sizeO < 0 || sizeP < 0 || nF <= 0) return fail(NDA_ERR_ARG, "bad size");
    if (maxN < 1 || maxN > SURF_MAXN) return fail(NDA_ERR_ARG, "maxN = %d outside [1, %d]", maxN, SURF_MAXN);
    if (sizeO == 0) return NDA_OK;
    NDA_ENTER();
    cudaStream_t st = c->stream;
    Ctx::Lane &L = c->lane[0];
    const size_t wBytes = (size_t)sizeO * nF * 12;
    CK(c->raw[0][0].reserve(wBytes + 16));
    CK(c->raw[1][0].reserve(wBytes + 16));
    CK(cudaMemcpyAsync(c->raw[0][0].p, waterO, wBytes, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->raw[1][0].p, watVec, wBytes, cudaMemcpyHostToDevice, st));
    float *d_prot = nullptr, *d_cell = nullptr;
    rc = surface_common(*c, st, prot, sizeP, cellDims, nF, &d_prot, &d_cell);
    if (rc) return rc;
    const int nPPad = round_up(std::max(sizeP, 1), 32);
    const int pass = std::min(32768, frames_per_pass((size_t)nPPad * sizeof(float4), nF));
    CK(L.packB.reserve((size_t)nPPad * pass * sizeof(float4)));
    CK(L.maxAbs.reserve((size_t)pass * sizeof(unsigned)));
    for (int f0 = 0; f0 < nF; f0 += pass) {
        const int nFc = std::min(pass, nF - f0);
        CK(cudaMemsetAsync(L.maxAbs.p, 0, (size_t)nFc * sizeof(unsigned), st));
        rc = launch_pack(*c, st, d_prot, nF, nullptr, sizeP, nPPad, f0, nFc, nullptr, L.packB.as<float4>(),
                         L.maxAbs.as<unsigned>(), d_cell, nullptr);
        if (rc) return rc;
        dim3 grid((sizeO + SURF_THREADS - 1) / SURF_THREADS, nFc);
        CK(c->kev_record(st));
        water_orient_kernel<<<grid, SURF_THREADS, 0, st>>>(c->raw[0][0].as<float>(), c->raw[1][0].as<float>(), sizeO, nF, f0,
                                                           L.packB.as<float4>(), sizeP, nPPad, d_cell, minR, maxR, maxN);
        LAUNCHED();
        CK(c->kev_record(st));
    }
    CK(cudaMemcpyAsync(waterO, c->raw[0][0].p, wBytes, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return NDA_OK;
}

// ---- hydrogen-bond autocorrelation ----------------------------------------------
int nda_get_hb_counts(const float *acceptors, int nA, int nF, const float *donors, int nD,
                      const float *hydrogens, int nH, const float *cellDims, uint64_t *counts,
                      float maxR, float minAngle, int continuous)
{
    if (!acceptors || !donors || !hydrogens || !cellDims || !counts) return fail(NDA_ERR_ARG, "null pointer argument");
    if (nA < 0 || nD < 0 || nH < nD || nF <= 0) return fail(NDA_ERR_ARG, "bad size (every donor needs its hydrogen)");
    for (int dt = 0; dt < nF; ++dt) counts[dt] = 0;
    if (nA == 0 || nD == 0) return NDA_OK;
    NDA_ENTER();
    cudaStream_t st = c->stream;
    const float cosAngle = cosf(minAngle);                  // getHydrogenBonds.cpp:16, the host's libm like the reference
    CK(c->raw[0][0].reserve((size_t)nA * nF * 12 + 16));
    CK(c->raw[0][1].reserve((size_t)nD * nF * 12 + 16));
    CK(c->raw[1][0].reserve((size_t)nD * nF * 12 + 16));
    CK(c->cell.reserve((size_t)nF * 12));
    CK(c->counts.reserve((size_t)nF * 8 + 16));
    CK(cudaMemcpyAsync(c->raw[0][0].p, acceptors, (size_t)nA * nF * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->raw[0][1].p, donors, (size_t)nD * nF * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->raw[1][0].p, hydrogens, (size_t)nD * nF * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(c->cell.p, cellDims, (size_t)nF * 12, cudaMemcpyHostToDevice, st));
    CK(cudaMemsetAsync(c->counts.p, 0, (size_t)nF * 8, st));
    unsigned *d_n = c->lane[0].ctl.as<unsigned>() + 2;      // spare word of the control block
    unsigned capacity = 64u * (unsigned)std::max(nA, nD) + 4096u;
    unsigned nPairs = 0;
    for (int attempt = 0; attempt < 2; ++attempt) {
        CK(c->idx1.reserve((size_t)capacity * 8));
        CK(cudaMemsetAsync(d_n, 0, 4, st));
        dim3 grid((nD + 255) / 256, std::min(nA, 4096));
        hb_frame0_kernel<<<grid, 256, 0, st>>>(c->raw[0][0].as<float>(), nA, c->raw[0][1].as<float>(), nD,
                                               c->raw[1][0].as<float>(), c->cell.as<float>(), nF, maxR, cosAngle,
                                               c->idx1.as<int>(), capacity, d_n);
        LAUNCHED();
        CK(cudaMemcpyAsync(&nPairs, d_n, 4, cudaMemcpyDeviceToHost, st));
        CK(cudaStreamSynchronize(st));
        if (nPairs <= capacity) break;
        capacity = nPairs;                                   // denser than guessed: redo with the exact size
    }
    if (nPairs > 0) {
        CK(c->kev_record(st));
        hb_walk_kernel<<<(nPairs + 7) / 8, 256, 0, st>>>(c->raw[0][0].as<float>(), c->raw[0][1].as<float>(),
                                                         c->raw[1][0].as<float>(), c->cell.as<float>(), nF, maxR, cosAngle,
                                                         continuous == 1 ? 1 : 0, c->idx1.as<int>(), nPairs,
                                                         c->counts.as<unsigned long long>());
        LAUNCHED();
        CK(c->kev_record(st));
    }
    CK(cudaMemcpyAsync(counts, c->counts.p, (size_t)nF * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    c->statsPending = false;
    c->stats[0] = 0.5 * (double)nA * (double)nD;
    c->stats[1] = (double)nPairs;
    return NDA_OK;
}

int nda_get_hb_corr(const float *acceptors, int nA, int nF, const float *donors, int nD,
                    const float *hydrogens, int nH, const float *cellDims, float *out,
                    int maxTime, int step, int nbrTimeOri, float maxR, float minAngle, int continuous)
{
    (void)maxTime; (void)step; (void)nbrTimeOri;            // unused by the reference as well
    if (!out) return fail(NDA_ERR_ARG, "null pointer argument");
    if (nF <= 0) return fail(NDA_ERR_ARG, "bad size");
    std::vector<uint64_t> cnt((size_t)nF, 0);
    int rc = nda_get_hb_counts(acceptors, nA, nF, donors, nD, hydrogens, nH, cellDims, cnt.data(), maxR, minAngle, continuous);
    if (rc) return rc;
    for (int dt = 0; dt < nF; ++dt) out[dt] += (float)cnt[dt];   // the reference ACCUMULATES into out (:73)
    return NDA_OK;
}

// ---- DCD ingest ------------------------------------------------------------------
// Drop-ins for lib/common/src/dcdReader.cpp:28-98 (getDCDCoor) and dcdCellReader.cpp:26-70
// (getDCDCell), which do not compile outside MSVC (_fseeki64, undeclared lastAtomIdx) and byte-swap
// nothing.  Same arguments, same error codes (0, -1 open, -2 seek/out of range, -3 short read), same
// output layout; pread() instead of a shared FILE* so the frames are read by several host threads,
// and a real byte swap when the file's endianness differs from the host's.
static inline uint32_t bswap32(uint32_t v) { return __builtin_bswap32(v); }

int nda_dcd_get_coor(const char *fileName, const int *frames, int nbrFrames, int nbrAtoms,
                     const int *selAtoms, int selAtomsSize, const int *dims, int nbrDims, int cell,
                     const long long *startPos, float *outArr, char byteorder)
{
    if (!fileName || !frames || !selAtoms || !dims || !startPos || !outArr) return -1;
    if (nbrAtoms <= 0 || nbrFrames < 0 || selAtomsSize < 0 || nbrDims < 0) return -2;
    for (int i = 0; i < selAtomsSize; ++i) if (selAtoms[i] < 0 || selAtoms[i] >= nbrAtoms) return -2;
    for (int d = 0; d < nbrDims; ++d) if (dims[d] < 0 || dims[d] > 2) return -2;
    const int fd = open(fileName, O_RDONLY);
    if (fd < 0) return -1;
    const uint16_t probe = 1;
    const char sys = *reinterpret_cast<const char *>(&probe) ? '<' : '>';
    const bool swap = (byteorder == '<' || byteorder == '>') && byteorder != sys;
    int err = 0;
    #pragma omp parallel num_threads(default_host_threads())
    {
        std::vector<uint32_t> rec((size_t)nbrAtoms);
        #pragma omp for schedule(dynamic, 1)
        for (int frameId = 0; frameId < nbrFrames; ++frameId) {
            const int frame = frames[frameId];
            for (int dimId = 0; dimId < nbrDims; ++dimId) {
                // dcdReader.cpp:66-68: record of axis dims[dimId] of this frame
                const long long pos = startPos[frame] + (long long)dims[dimId] * (4LL * nbrAtoms + 8) + (cell ? 60 : 4);
                size_t got = 0;
                const size_t want = 4u * (size_t)nbrAtoms;
                while (got < want) {
                    const ssize_t r = pread(fd, reinterpret_cast<char *>(rec.data()) + got, want - got, pos + (long long)got);
                    if (r <= 0) break;
                    got += (size_t)r;
                }
                if (got < want) {
                    #pragma omp atomic write
                    err = (pos < 0) ? -2 : -3;
                    continue;
                }
                for (int a = 0; a < selAtomsSize; ++a) {        // dcdReader.cpp:88-89
                    uint32_t v = rec[selAtoms[a]];
                    if (swap) v = bswap32(v);
                    memcpy(&outArr[(size_t)nbrDims * nbrFrames * a + (size_t)nbrDims * frameId + dimId], &v, 4);
                }
            }
        }
    }
    close(fd);
    return err;
}

int nda_dcd_get_cell(const char *fileName, const int *frames, int nbrFrames, const long long *startPos,
                     double *outArr, char byteorder)
{
    if (!fileName || !frames || !startPos || !outArr) return -1;
    const int fd = open(fileName, O_RDONLY);
    if (fd < 0) return -1;
    const uint16_t probe = 1;
    const char sys = *reinterpret_cast<const char *>(&probe) ? '<' : '>';
    const bool swap = (byteorder == '<' || byteorder == '>') && byteorder != sys;
    int err = 0;
    for (int frameId = 0; frameId < nbrFrames; ++frameId) {
        uint64_t rec[6];
        const ssize_t r = pread(fd, rec, sizeof rec, startPos[frames[frameId]] + 4);   // dcdCellReader.cpp:44
        if (r != (ssize_t)sizeof rec) { err = -2; break; }
        for (int i = 0; i < 6; ++i) {
            uint64_t v = swap ? __builtin_bswap64(rec[i]) : rec[i];
            memcpy(&outArr[6 * (size_t)frameId + i], &v, 8);
        }
    }
    close(fd);
    return err;
}

// ---- measurement / test support --------------------------------------------------------------
int nda_dev_synth_water_box(float *d_xyz, float *d_cellDims, int n_side, int nF, int fb, int fe,
                            float L, float sigma, uint32_t seed, void *stream)
{
    if (!d_xyz || !d_cellDims || n_side <= 0 || n_side > 1024 || fb < 0 || fe > nF || fb > fe)
        return fail(NDA_ERR_ARG, "bad argument");
    if (fe == fb) return NDA_OK;
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const unsigned n = (unsigned)n_side * n_side * n_side;
    for (int f0 = fb; f0 < fe; f0 += 32768 * SYNTH_FB) {
        const int nFc = std::min(32768 * SYNTH_FB, fe - f0);
        dim3 grid((n + 255) / 256, (nFc + SYNTH_FB - 1) / SYNTH_FB);
        synth_water_box_kernel<<<grid, 256, 0, st>>>(d_xyz, d_cellDims, n_side, nF, f0, nFc, L, sigma, seed);
        LAUNCHED();
    }
    return NDA_OK;
}

// frames [firstFrame, firstFrame + nFrames) of the same generator into columns [0, nFrames) of a LOCAL array with a row
// stride of strideFrames frames: a rank that owns a frame range of a long trajectory generates only its own frames
int nda_dev_synth_water_box_slice(float *d_xyz, float *d_cellDims, int n_side, int strideFrames, int firstFrame, int nFrames,
                                  float L, float sigma, uint32_t seed, void *stream)
{
    if (!d_xyz || !d_cellDims || n_side <= 0 || n_side > 1024 || firstFrame < 0 || nFrames < 0 || nFrames > strideFrames)
        return fail(NDA_ERR_ARG, "bad argument");
    if (nFrames == 0) return NDA_OK;
    NDA_ENTER();
    cudaStream_t st = stream ? (cudaStream_t)stream : c->stream;
    const unsigned n = (unsigned)n_side * n_side * n_side;
    // the kernel writes frame f at column f: point it firstFrame columns before the local array
    float *xyz0 = d_xyz - (ptrdiff_t)firstFrame * 3, *cell0 = d_cellDims - (ptrdiff_t)firstFrame * 3;
    for (int f0 = firstFrame; f0 < firstFrame + nFrames; f0 += 32768 * SYNTH_FB) {
        const int nFc = std::min(32768 * SYNTH_FB, firstFrame + nFrames - f0);
        dim3 grid((n + 255) / 256, (nFc + SYNTH_FB - 1) / SYNTH_FB);
        synth_water_box_kernel<<<grid, 256, 0, st>>>(xyz0, cell0, n_side, strideFrames, f0, nFc, L, sigma, seed);
        LAUNCHED();
    }
    return NDA_OK;
}

int nda_measure_fp32_peak(int variant, double *tflops)
{
    if (!tflops) return fail(NDA_ERR_ARG, "null pointer argument");
    NDA_ENTER();
    cudaStream_t st = c->stream;
    CK(c->bits.reserve(64));
    const int iters = 4096, blocks = c->numSMs * 8;
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaEventRecord(c->ev0, st));
        if (variant == 0) fp32_peak_kernel<0><<<blocks, 256, 0, st>>>(c->bits.as<float>(), iters, 0.999f, 0.001f);
        else              fp32_peak_kernel<1><<<blocks, 256, 0, st>>>(c->bits.as<float>(), iters, 0.999f, 0.001f);
        LAUNCHED();
        CK(cudaEventRecord(c->ev1, st));
        CK(cudaStreamSynchronize(st));
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
        const double flop = 2.0 * 16.0 * iters * 256.0 * blocks;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    return NDA_OK;
}

// Page-lock a caller's array (cudaHostRegister) so the host entry points DMA it directly -- strided 2-D copies
// straight from the array, no gather through staging -- which is what lets several devices pull their frame
// columns at link rate at the same time.  The caller keeps the array alive until nda_host_unregister.
int nda_host_register(const void *ptr, size_t bytes)
{
    if (!ptr || bytes == 0) return fail(NDA_ERR_ARG, "null pointer / empty range");
    Ctx *c;
    int rc = ctx_get(&c);
    if (rc) return rc;
    cudaError_t e = cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterPortable | cudaHostRegisterReadOnly);
    if (e != cudaSuccess) {                                 // read-only registration is not available on every platform
        cudaGetLastError();
        e = cudaHostRegister(const_cast<void *>(ptr), bytes, cudaHostRegisterPortable);
    }
    if (e != cudaSuccess) return fail(NDA_ERR_CUDA, "cudaHostRegister of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return NDA_OK;
}

int nda_host_unregister(const void *ptr)
{
    if (!ptr) return fail(NDA_ERR_ARG, "null pointer");
    CK(cudaHostUnregister(const_cast<void *>(ptr)));
    return NDA_OK;
}

// One production tcgen05.mma on caller-supplied fp16 operands (A [128][16], B [256][16], raw bits) and the
// production packed read-back: out uint32 [128][128], out[row][cq * 32 + n] = register n of columns 64 cq ..
int nda_debug_tc_probe(const uint16_t *A, const uint16_t *B, uint32_t *out)
{
    if (!A || !B || !out) return fail(NDA_ERR_ARG, "null pointer argument");
    NDA_ENTER();
    cudaStream_t st = c->stream;
    CK(c->raw[0][0].reserve(TC_SBLK * 32 + TC_RBLK * 32));
    CK(c->bits.reserve(128 * 128 * 4));
    __half *dA = c->raw[0][0].as<__half>(), *dB = dA + TC_SBLK * 16;
    CK(cudaMemcpyAsync(dA, A, TC_SBLK * 32, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(dB, B, TC_RBLK * 32, cudaMemcpyHostToDevice, st));
    const size_t smem = TC_SEMB_BYTES + TC_REMB_BYTES + 64;
    CK(cudaFuncSetAttribute((const void *)tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_probe_kernel<<<1, 512, smem, st>>>(dA, dB, c->bits.as<uint32_t>());
    LAUNCHED();
    CK(cudaMemcpyAsync(out, c->bits.p, 128 * 128 * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return NDA_OK;
}

// Platform ceiling of the end-to-end path: `bytes` of CONTIGUOUS pinned host memory copied to every device of
// the calling thread at the same time (one plain cudaMemcpyAsync each, one host thread per device), `reps`
// times; returns the aggregate GB/s of the slowest repetition-averaged device (max over devices of the time).
int nda_measure_h2d(size_t bytes, int reps, double *gbps)
{
    if (!gbps || bytes == 0 || reps < 1) return fail(NDA_ERR_ARG, "bad argument");
    std::vector<Shard> sh;
    int rc = plan_shards(0, 1 << 20, sh);
    if (rc) return rc;
    const int G = (int)sh.size();
    std::vector<double> ms((size_t)G, 0.0);
    HostBarrier bar(G);
    rc = run_shards(sh, [&](int g, Ctx &c) -> int {
        void *h = nullptr;
        cudaError_t e = cudaMallocHost(&h, bytes);
        if (e == cudaSuccess) e = c.raw[0][0].reserve(bytes);
        if (e == cudaSuccess) { memset(h, 1, bytes); e = cudaMemcpyAsync(c.raw[0][0].p, h, bytes, cudaMemcpyHostToDevice, c.copyStream); }
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.copyStream);
        bar.wait();                                          // every device warmed up: start together
        if (e == cudaSuccess) e = cudaEventRecord(c.ev0, c.copyStream);
        for (int r = 0; r < reps && e == cudaSuccess; ++r)
            e = cudaMemcpyAsync(c.raw[0][0].p, h, bytes, cudaMemcpyHostToDevice, c.copyStream);
        if (e == cudaSuccess) e = cudaEventRecord(c.ev1, c.copyStream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c.copyStream);
        float t = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&t, c.ev0, c.ev1);
        ms[g] = t;
        if (h) cudaFreeHost(h);
        if (e != cudaSuccess) return fail(NDA_ERR_CUDA, "H2D measurement failed: %s", cudaGetErrorString(e));
        return NDA_OK;
    });
    if (rc) return rc;
    double worst = 0.0;
    for (double t : ms) worst = std::max(worst, t);
    *gbps = worst > 0.0 ? (double)bytes * reps * G / (worst * 1e-3) / 1e9 : 0.0;
    return NDA_OK;
}

} // extern "C"
