/* x265cu.cu -- host side of libx265cu.so: context, device mirrors of `Lowres`, batching and the
 * extern "C" entry points declared in include/x265cu.h.  No CPU fallback anywhere: every entry
 * either runs its CUDA kernels or fails with an error code.
 *
 * HBM layout (one allocation per array family, sized at x265cu_open for numFrameSlots mirrors):
 *   planes      [slot][4][planeSize]                padded lowres planes, layout of Lowres::buffer[0]
 *   intraCost   [slot][nCU] i32      intraMode [slot][nCU] u8      invQ [slot][nCU] i32
 *   lowresCosts [slot][bf+2][bf+2][nCU] u16         rowSatds [slot][bf+2][bf+2][hCU] i32
 *   mvs         [slot][2][bf+1][nCU] (int16 x, int16 y)      mvCosts [slot][2][bf+1][nCU] i32
 *   propagate   [slot][nCU] u64                     cuTree accumulators (x265cu_cutree.cuh)
 *   wplanes     pool of weighted 4-plane copies (one per weighted job of a batch)
 * Results of a batch are also written to one packed device record per job.  From there each array goes straight
 * into the caller's Lowres array when that lies in mapped pinned memory (x265cu_host_register: scatter_results_kernel
 * writes it over PCIe), otherwise through one device->host copy into pinned staging and a host memcpy; the 32-byte
 * sums of every job always come back by copy.
 * Streams: the ctx's compute stream; an upload stream (pictures of a pre-lookahead list / x265cu_frame_upload); an
 * intra stream (intra estimates of a list beside the uploads of its later frames); a copy stream (plane copy-backs).
 */
#include "../../../include/x265cu.h"
#include "x265cu_kernels.cuh"
#include "x265cu_cutree.cuh"

#include <cuda_runtime.h>
#include <sched.h>
#include <stdio.h>
#include <time.h>
#include <unistd.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cuda.h>
#include <map>
#include <mutex>
#include <thread>
#include <vector>

namespace {

char g_openError[512] = "";

/* host ranges registered through x265cu_host_register, with their device-side alias (mapped pinned memory): result
 * arrays whose destination lies in one of them are written there by a kernel instead of being staged and memcpy'd */
struct HostRange { uintptr_t base; size_t bytes; uintptr_t dev; };
std::mutex g_regMtx;
std::map<uintptr_t, HostRange> g_reg;

uint8_t* mappedAlias(const void* p, size_t bytes)
{
    std::lock_guard<std::mutex> lk(g_regMtx);
    std::map<uintptr_t, HostRange>::const_iterator it = g_reg.upper_bound((uintptr_t)p);
    if (it == g_reg.begin()) return NULL;
    --it;
    const HostRange& r = it->second;
    if ((uintptr_t)p < r.base || (uintptr_t)p + bytes > r.base + r.bytes) return NULL;
    return (uint8_t*)(r.dev + ((uintptr_t)p - r.base));
}

struct PendingEvent { int kind; cudaEvent_t a, b; };

} // namespace

struct x265cu_ctx
{
    x265cu_config cfg;
    GeomDev g;
    int pb;            /* bytes per sample */
    int pixelMax;
    int correction;    /* IF_INTERNAL_PREC - X265_DEPTH */
    int bf;
    cudaStream_t stream;
    bool ownStream;
    cudaStream_t copyStream;               /* device->host copies of the padded planes run behind the compute stream */
    std::vector<cudaEvent_t> planesCopied; /* per slot: last planes copy-back finished */
    std::vector<char> planesPending;
    /* copy-backs of a pre-lookahead LIST are held back until the next estimate batch is running (or x265cu_sync):
     * while the list's pictures are being uploaded, PCIe is the bottleneck and a copy the other way slows it down;
     * while the estimates compute, the bus is idle */
    std::vector<cudaEvent_t> planesReady;  /* per slot: the lowres kernel that produced the planes finished */
    struct DeferredPlanes { int slot; void* dst; };
    std::vector<DeferredPlanes> deferredPlanes;
    bool deferPlanes;
    cudaEvent_t evKernel;
    std::mutex mtx;
    char err[512];

    uint8_t* dPlanes;
    int* dIntraCost;
    uint8_t* dIntraMode;
    int* dInvQ;
    std::vector<char> hasInvQ;
    uint16_t* dLowresCosts;
    int* dRowSatds;
    int* dMvs;
    int* dMvCosts;
    unsigned long long* dPropagate;        /* [slot][nCU] Lowres::propagateCost accumulators (x265cu_cutree.cuh) */
    /* explicit weighted-prediction analysis (x265cu_wp.cuh): compact copies of the source chroma planes per slot, the
     * motion-compensated reference plane and the vectors / intra costs of the (slice, list, plane) being analysed */
    std::vector<std::pair<void*, size_t> > fixedAllocs;     /* arrays of x265cu_open, returned to the pool by x265cu_close */
    uint8_t* dChroma; int chromaPitch, chromaRows; std::vector<char> hasChroma;
    uint8_t* dWp; size_t dWpCap;
    WpCostArgs wpArgs; bool wpReady;
    int cutreeCtas;                        /* grid of the cooperative cuTree kernel (one CTA per SM) */
    const char* lastEnqueued;              /* diagnostics (X265CU_SLOW_LOG_MS): the entry that last put work on the stream */
    bool mappedResults;                    /* result arrays go straight into mapped pinned destinations (X265CU_MAPPED_RESULTS=0: always staged) */
    uint16_t* dPropOut; size_t dPropOutCap; /* clamped uint16 copies on their way to the host */
    uint16_t* hPropOut; size_t hPropOutCap;
    uint16_t* dLut;        /* base; centre at +65536 */
    uint8_t* dSrc;         /* full-resolution luma staging (strided-copy fallback) */
    uint8_t* dSrcLin; size_t dSrcLinCap;   /* luma staging with the host's pitch (linear transfer) */
    uint8_t* dUp; size_t dUpCap;           /* per-frame staging of a batched pre-lookahead list */
    cudaStream_t upStream;                 /* its uploads */
    cudaStream_t intraStream;              /* x265cu_pre_lookahead_batch: intra estimates run beside the uploads of later frames */
    /* x265cu_frame_upload: a picture uploaded ahead of its pre-lookahead, per slot */
    struct SlotUpload { const void* y; const void* u; const void* v; intptr_t ys, cs; uint8_t* d; size_t cap; cudaEvent_t done, read; bool valid, everRead; };
    std::vector<SlotUpload> slotUp;
    std::vector<cudaEvent_t> preEvents;    /* per frame of the list: lowres planes + energies/sums done */
    std::vector<cudaEvent_t> upEvents;
    int64_t srcPitch;      /* samples */
    unsigned long long* dSmall;   /* small scratch for sums */

    /* growable staging */
    uint8_t* dStage; size_t dStageCap;
    uint8_t* hStage; size_t hStageCap;
    uint8_t* dArgs; size_t dArgsCap;       /* JobDev[], SearchItem[], int[] */
    uint8_t* hArgs; size_t hArgsCap;
    uint8_t* hPre; size_t hPreCap;         /* pinned landing zone of batched pre-lookahead results */
    uint8_t* dPre; size_t dPreCap;         /* its device mirror */
    std::vector<void*> wPool;              /* weighted plane sets */
    uint8_t* dGeneric; size_t dGenericCap; /* pixelcmp / var scratch */
    uint8_t* dMemo; size_t dMemoCap;       /* search memo of a batch: [search][nCU][MEMO_N] int4 */
    /* hint bookkeeping for the speculation kernel (affects speed only, never a result): relative
     * picture order of the frame slots inferred from the jobs (poc(fenc) - poc(ref0) = d0, ...) and
     * which MV mirrors hold a finished search */
    std::vector<long long> slotPoc;
    long long pocBase;
    std::vector<char> slotPocKnown;
    std::vector<char> mvValid;             /* [slot][list][d - 1] */

    bool timing;
    std::vector<PendingEvent> pending;
    std::vector<cudaEvent_t> freeEvents;
    x265cu_stats stats;

    int searchWarps;       /* CU rows (= warps) per commit CTA */
    int searchSpec;        /* refine iterations before the commit wavefront (0: none; experiments only) */
    int searchMode;        /* 0: plain kernel only, 1: speculative path only (with in-batch seed waves), 2: per search (default) */
    int specMaxDist, specMaxPlans;   /* speculative path: hint at most this many frames away, batch of at most this many searches */
    int plainWarps;        /* CU rows (= warps) per CTA of the plain kernel */
    int plainTma;          /* plain kernel's windows staged by TMA (cp.async.bulk.tensor.3d), the next CU's window requested a step ahead */
    bool tmaMapValid;
    PlainTmaMap tmaMap;    /* the frame mirrors as a tensor [slot x 4 planes][padded rows][stride] */
    int plainOneShot;      /* plain kernel with a window: the no-move positions of a search measured in one burst (la_fast_path) */
    int plainOct;          /* plain wavefront search: 1 = octet kernel (x265cu_search_oct.cuh, default), 0 = quad kernel (x265cu_search_plain.cuh) */
    int octWarps;          /* octet kernel: bands of 4 CU rows (= warps) per CTA */
    int octSlack, octSleep, octSleepFull, octFullWarps;   /* hand-off waits: extra distance (columns) and poll interval (ns) of launches with >= octFullWarps warps */
    int plainWin;          /* plain kernel's shared-memory window variant: -2 for 8-bit samples (default; measured 5 % slower at 16 bit), 1 always, 0 never, -1 only for launches that fill the GPU */
    long long dbgPlans[4];
    double hostMs[4], planMs[4];      /* estimate_batch host time: planning, enqueue, wait, scatter (X265CU_HOST_PROFILE) */
    long long hostCalls;
};

namespace {

/* host threads for the memcpy scatter of a big batch (X265CU_HOST_THREADS overrides; several ranks share the cores) */
unsigned hostThreads()
{
    if (const char* e = getenv("X265CU_HOST_THREADS")) return (unsigned)atoi(e);
    /* the cores this process may run on (a pinned rank sees its share, not the whole box) */
    cpu_set_t set;
    unsigned n = std::thread::hardware_concurrency();
    if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = (unsigned)CPU_COUNT(&set);
    return n >= 8 ? n / 4 : (n >= 2 ? 2 : 1);
}

#define CU_TRY(ctx, call)                                                                           \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
        {                                                                                           \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return X265CU_ECUDA;                                                                    \
        }                                                                                           \
    } while (0)

int fail(x265cu_ctx* c, int code, const char* msg)
{
    snprintf(c->err, sizeof(c->err), "%s", msg);
    return code;
}

inline size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* Work buffers (argument / staging / memo areas) grow on demand.  cudaMalloc / cudaMallocHost / cudaFree cost 0.1-10 ms
 * each and synchronise the device, so a buffer a context no longer needs (it grew, or the context closed) goes to a
 * process-wide pool per device and the next request of any context is served from there: a new encoder instance in a
 * process that has run one before starts warm.  x265cu_trim() empties the pool. */
struct PooledBuf { void* p; size_t cap; int device; };
std::mutex g_poolMtx;
std::vector<PooledBuf> g_devPool, g_hostPool;

void* poolTake(std::vector<PooledBuf>& pool, int device, size_t need, size_t* cap)
{
    std::lock_guard<std::mutex> lk(g_poolMtx);
    int best = -1;
    for (size_t i = 0; i < pool.size(); i++)
        if (pool[i].device == device && pool[i].cap >= need && (best < 0 || pool[i].cap < pool[(size_t)best].cap)) best = (int)i;
    if (best < 0) return NULL;
    void* p = pool[(size_t)best].p;
    *cap = pool[(size_t)best].cap;
    pool[(size_t)best] = pool.back();
    pool.pop_back();
    return p;
}

void poolGive(std::vector<PooledBuf>& pool, int device, void* p, size_t cap)
{
    if (!p) return;
    std::lock_guard<std::mutex> lk(g_poolMtx);
    PooledBuf b = { p, cap, device };
    pool.push_back(b);
}

/* the per-context arrays of x265cu_open: same sizes for every context of a configuration, so a closed context's arrays serve
 * the next one (an encoder farm opens and closes contexts all the time; cudaMalloc / cudaFree stall every stream of the device) */
cudaError_t fixedAlloc(x265cu_ctx* c, void** p, size_t bytes)
{
    size_t cap = 0;
    void* q = poolTake(g_devPool, c->cfg.device, bytes, &cap);
    if (q && cap != bytes) { poolGive(g_devPool, c->cfg.device, q, cap); q = NULL; }
    if (!q)
    {
        cudaError_t e = cudaMalloc(&q, bytes);
        if (e != cudaSuccess) return e;
    }
    *p = q;
    c->fixedAllocs.push_back(std::make_pair(q, bytes));
    return cudaSuccess;
}

int growDevice(x265cu_ctx* c, uint8_t** p, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    poolGive(g_devPool, c->cfg.device, *p, *cap);
    *p = NULL; *cap = 0;
    size_t n = alignUp(need + need / 2, 1 << 20);
    if ((*p = (uint8_t*)poolTake(g_devPool, c->cfg.device, need, cap)) != NULL) return 0;
    CU_TRY(c, cudaMalloc((void**)p, n));
    *cap = n;
    return 0;
}

int growHost(x265cu_ctx* c, uint8_t** p, size_t* cap, size_t need)
{
    if (*cap >= need) return 0;
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    poolGive(g_hostPool, c->cfg.device, *p, *cap);
    *p = NULL; *cap = 0;
    size_t n = alignUp(need + need / 2, 1 << 20);
    if ((*p = (uint8_t*)poolTake(g_hostPool, c->cfg.device, need, cap)) != NULL) return 0;
    CU_TRY(c, cudaMallocHost((void**)p, n));
    *cap = n;
    return 0;
}

/* ---- streams and events are pooled across the contexts of a process too (per device).  A host that opens a context per
 * encode (or, like the lookahead-only driver, per run) would otherwise create and destroy a handful of streams and a few
 * hundred events each time; stream creation / destruction goes through the driver's system-wide resource manager, and with
 * one process per GPU doing that at the same moment on an 8-GPU box the calls of OTHER processes stall behind it (measured:
 * cudaStreamSynchronize of an idle stream blocked for 0.6-1.4 s).  A stream is given back idle (x265cu_close waits first). ---- */
struct ResPool { std::vector<cudaStream_t> streams; std::vector<cudaEvent_t> events, timingEvents; };
std::map<int, ResPool> g_resPool;
std::mutex g_resMtx;

cudaError_t takeStream(int dev, cudaStream_t* s)
{
    {
        std::lock_guard<std::mutex> lk(g_resMtx);
        ResPool& rp = g_resPool[dev];
        if (!rp.streams.empty()) { *s = rp.streams.back(); rp.streams.pop_back(); return cudaSuccess; }
    }
    return cudaStreamCreateWithFlags(s, cudaStreamNonBlocking);
}
void giveStream(int dev, cudaStream_t s)
{
    if (!s) return;
    std::lock_guard<std::mutex> lk(g_resMtx);
    g_resPool[dev].streams.push_back(s);
}
cudaError_t takeEvent(int dev, cudaEvent_t* e, bool timing = false)
{
    {
        std::lock_guard<std::mutex> lk(g_resMtx);
        ResPool& rp = g_resPool[dev];
        std::vector<cudaEvent_t>& v = timing ? rp.timingEvents : rp.events;
        if (!v.empty()) { *e = v.back(); v.pop_back(); return cudaSuccess; }
    }
    return timing ? cudaEventCreate(e) : cudaEventCreateWithFlags(e, cudaEventDisableTiming);
}
void giveEvent(int dev, cudaEvent_t e, bool timing = false)
{
    if (!e) return;
    std::lock_guard<std::mutex> lk(g_resMtx);
    ResPool& rp = g_resPool[dev];
    (timing ? rp.timingEvents : rp.events).push_back(e);
}

/* entries without a context (pinned-memory registration) run on the device of the most recent x265cu_open of the process, else on
 * $X265CU_DEVICE, never on the runtime's default device 0 by accident (a process per GPU must not create a context on GPU 0) */
std::atomic<int> g_defaultDevice(-1);
void bindDefaultDevice()
{
    int dev = g_defaultDevice.load();
    if (dev < 0 && getenv("X265CU_DEVICE")) dev = atoi(getenv("X265CU_DEVICE"));
    if (dev >= 0) { if (cudaSetDevice(dev) != cudaSuccess) cudaGetLastError(); }
}

/* ---- timing helpers ---- */
cudaEvent_t getEvent(x265cu_ctx* c)
{
    if (!c->freeEvents.empty()) { cudaEvent_t e = c->freeEvents.back(); c->freeEvents.pop_back(); return e; }
    cudaEvent_t e = NULL;
    takeEvent(c->cfg.device, &e, true);
    return e;
}

struct KernelScope
{
    x265cu_ctx* c; int kind; cudaEvent_t a, b; bool on; cudaStream_t st;
    KernelScope(x265cu_ctx* ctx, int k, int launches = 1, cudaStream_t stream = NULL) : c(ctx), kind(k), on(ctx->timing), st(stream ? stream : ctx->stream)
    {
        c->stats.launches[kind] += launches;
        if (on) { a = getEvent(c); b = getEvent(c); cudaEventRecord(a, st); }
    }
    ~KernelScope()
    {
        if (on) { cudaEventRecord(b, st); PendingEvent p = { kind, a, b }; c->pending.push_back(p); }
    }
};

void resolveEvents(x265cu_ctx* c)
{
    for (size_t i = 0; i < c->pending.size(); i++)
    {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->pending[i].a, c->pending[i].b) == cudaSuccess)
            c->stats.ms[c->pending[i].kind] += ms;
        c->freeEvents.push_back(c->pending[i].a);
        c->freeEvents.push_back(c->pending[i].b);
    }
    c->pending.clear();
}

/* X265CU_SLOW_LOG_MS=N (diagnostics): every library call / wait that takes longer than N ms is reported on stderr with a wall-clock
 * stamp, so that stalls of several processes sharing a box can be lined up */
struct SlowLog
{
    const char* name; std::chrono::steady_clock::time_point t0; double limit;
    static double threshold() { static const double v = getenv("X265CU_SLOW_LOG_MS") ? atof(getenv("X265CU_SLOW_LOG_MS")) : 0.0; return v; }
    explicit SlowLog(const char* n) : name(n), limit(threshold()) { if (limit > 0) t0 = std::chrono::steady_clock::now(); }
    ~SlowLog()
    {
        if (limit <= 0) return;
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ms < limit) return;
        struct timespec ts;
        clock_gettime(CLOCK_REALTIME, &ts);
        fprintf(stderr, "x265cu slow call [pid %d, wall %ld.%03ld]: %s took %.1f ms\n", (int)getpid(), (long)(ts.tv_sec % 100000), ts.tv_nsec / 1000000, name, ms);
    }
};

int syncStream(x265cu_ctx* c)
{
    SlowLog slow("cudaStreamSynchronize(stream)");
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    resolveEvents(c);
    return 0;
}

/* issue the held-back plane copy-backs on the copy stream (each behind the kernel that produced its planes) */
int flushDeferredPlanes(x265cu_ctx* c)
{
    const size_t bytes = (size_t)4 * c->g.planeSize * c->pb;
    for (size_t k = 0; k < c->deferredPlanes.size(); k++)
    {
        const int slot = c->deferredPlanes[k].slot;
        CU_TRY(c, cudaStreamWaitEvent(c->copyStream, c->planesReady[slot], 0));
        CU_TRY(c, cudaMemcpyAsync(c->deferredPlanes[k].dst, c->dPlanes + (size_t)slot * 4 * c->g.planeSize * c->pb, bytes, cudaMemcpyDeviceToHost, c->copyStream));
        CU_TRY(c, cudaEventRecord(c->planesCopied[slot], c->copyStream));
        c->planesPending[slot] = 1;
        c->stats.d2hBytes += (int64_t)bytes;
    }
    c->deferredPlanes.clear();
    return 0;
}

/* ---- mirrors ---- */
inline uint8_t* slotBuffer(x265cu_ctx* c, int slot) { return c->dPlanes + (size_t)slot * 4 * c->g.planeSize * c->pb; }
inline uint8_t* slotPlane0(x265cu_ctx* c, int slot) { return slotBuffer(c, slot) + (size_t)c->g.padOffset * c->pb; }
inline int* slotIntraCost(x265cu_ctx* c, int slot) { return c->dIntraCost + (size_t)slot * c->g.nCU; }
inline uint8_t* slotIntraMode(x265cu_ctx* c, int slot) { return c->dIntraMode + (size_t)slot * c->g.nCU; }
inline int* slotInvQ(x265cu_ctx* c, int slot) { return c->dInvQ + (size_t)slot * c->g.nCU; }
inline uint16_t* slotLowresCosts(x265cu_ctx* c, int slot, int d0, int d1)
{
    return c->dLowresCosts + (((size_t)slot * (c->bf + 2) + d0) * (c->bf + 2) + d1) * c->g.nCU;
}
inline int* slotRowSatds(x265cu_ctx* c, int slot, int d0, int d1)
{
    return c->dRowSatds + (((size_t)slot * (c->bf + 2) + d0) * (c->bf + 2) + d1) * c->g.hCU;
}
inline int* slotMvs(x265cu_ctx* c, int slot, int list, int d) { return c->dMvs + (((size_t)slot * 2 + list) * (c->bf + 1) + (d - 1)) * c->g.nCU; }
inline int* slotMvCosts(x265cu_ctx* c, int slot, int list, int d) { return c->dMvCosts + (((size_t)slot * 2 + list) * (c->bf + 1) + (d - 1)) * c->g.nCU; }

inline uint8_t* slotChroma(x265cu_ctx* c, int slot, int plane) { return c->dChroma + ((size_t)slot * 2 + plane) * c->chromaPitch * c->chromaRows * c->pb; }

bool badSlot(const x265cu_ctx* c, int s) { return s < 0 || s >= c->cfg.numFrameSlots; }

void freeAll(x265cu_ctx* c)
{
    for (size_t i = 0; i < c->fixedAllocs.size(); i++) poolGive(g_devPool, c->cfg.device, c->fixedAllocs[i].first, c->fixedAllocs[i].second);
    c->fixedAllocs.clear();
    poolGive(g_devPool, c->cfg.device, c->dWp, c->dWpCap);
    /* the on-demand work buffers go back to the process-wide pool (growDevice / growHost); every stream of the context is
     * idle here (x265cu_close waited for them) */
    const int dev = c->cfg.device;
    poolGive(g_devPool, dev, c->dPropOut, c->dPropOutCap); poolGive(g_hostPool, dev, c->hPropOut, c->hPropOutCap);
    poolGive(g_devPool, dev, c->dStage, c->dStageCap); poolGive(g_devPool, dev, c->dArgs, c->dArgsCap);
    poolGive(g_devPool, dev, c->dGeneric, c->dGenericCap); poolGive(g_devPool, dev, c->dMemo, c->dMemoCap);
    poolGive(g_devPool, dev, c->dSrcLin, c->dSrcLinCap); poolGive(g_devPool, dev, c->dUp, c->dUpCap); poolGive(g_devPool, dev, c->dPre, c->dPreCap);
    poolGive(g_hostPool, dev, c->hStage, c->hStageCap); poolGive(g_hostPool, dev, c->hArgs, c->hArgsCap); poolGive(g_hostPool, dev, c->hPre, c->hPreCap);
    for (size_t i = 0; i < c->wPool.size(); i++) poolGive(g_devPool, dev, c->wPool[i], (size_t)4 * c->g.planeSize * c->pb + 256);
    giveStream(dev, c->upStream);
    giveStream(dev, c->intraStream);
    for (size_t i = 0; i < c->slotUp.size(); i++)
    {
        poolGive(g_devPool, dev, c->slotUp[i].d, c->slotUp[i].cap);
        giveEvent(dev, c->slotUp[i].done);
        giveEvent(dev, c->slotUp[i].read);
    }
    for (size_t i = 0; i < c->preEvents.size(); i++) giveEvent(dev, c->preEvents[i]);
    for (size_t i = 0; i < c->upEvents.size(); i++) giveEvent(dev, c->upEvents[i]);
    for (size_t i = 0; i < c->freeEvents.size(); i++) giveEvent(dev, c->freeEvents[i], true);
    if (c->ownStream) giveStream(dev, c->stream);
    giveStream(dev, c->copyStream);
    giveEvent(dev, c->evKernel);
    for (size_t i = 0; i < c->planesCopied.size(); i++) giveEvent(dev, c->planesCopied[i]);
    for (size_t i = 0; i < c->planesReady.size(); i++) giveEvent(dev, c->planesReady[i]);
}

} // namespace

/* ============================================================================================ */
extern "C" {

int x265cu_abi_version(void) { return X265CU_ABI_VERSION; }

int x265cu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char* x265cu_last_error(const x265cu_ctx* ctx) { return ctx ? ctx->err : g_openError; }

int x265cu_open(const x265cu_config* cfg, x265cu_ctx** out)
{
    SlowLog slow("x265cu_open");
    if (!cfg || !out) { snprintf(g_openError, sizeof(g_openError), "x265cu_open: NULL argument"); return X265CU_EINVAL; }
    *out = NULL;
    if (cfg->srcWidth < 16 || cfg->srcHeight < 16 || cfg->bframes < 0 || cfg->bframes > X265CU_BFRAME_MAX ||
        cfg->numFrameSlots < 1 || !cfg->mvcost || (cfg->bitDepth != 8 && cfg->bitDepth != 10 && cfg->bitDepth != 12) ||
        cfg->marginX < 32 || (cfg->marginX & 3) || cfg->marginY < 16 || cfg->numCoopSlices < 1 || cfg->numRowsPerSlice < 1)
    {
        snprintf(g_openError, sizeof(g_openError), "x265cu_open: unsupported configuration");
        return X265CU_EINVAL;
    }
    int ndev = x265cu_device_count();
    if (ndev <= 0 || cfg->device < 0 || cfg->device >= ndev)
    {
        snprintf(g_openError, sizeof(g_openError), "x265cu_open: no usable CUDA device (count=%d, requested %d); there is no CPU fallback", ndev, cfg->device);
        return X265CU_ENODEV;
    }
    x265cu_ctx* c = new x265cu_ctx();
    c->cfg = *cfg;
    c->err[0] = 0;
    c->dPlanes = NULL; c->dIntraCost = NULL; c->dIntraMode = NULL; c->dInvQ = NULL; c->dLowresCosts = NULL; c->dRowSatds = NULL;
    c->dMvs = NULL; c->dMvCosts = NULL; c->dLut = NULL; c->dSrc = NULL; c->dSmall = NULL;
    c->dStage = NULL; c->dStageCap = 0; c->hStage = NULL; c->hStageCap = 0; c->dArgs = NULL; c->dArgsCap = 0; c->hArgs = NULL; c->hArgsCap = 0; c->hPre = NULL; c->hPreCap = 0; c->dPre = NULL; c->dPreCap = 0; c->dSrcLin = NULL; c->dSrcLinCap = 0; c->dUp = NULL; c->dUpCap = 0; c->upStream = NULL; c->intraStream = NULL; c->lastEnqueued = NULL;
    c->dGeneric = NULL; c->dGenericCap = 0; c->dMemo = NULL; c->dMemoCap = 0;
    c->cutreeCtas = 0;
    c->mappedResults = !(getenv("X265CU_MAPPED_RESULTS") && atoi(getenv("X265CU_MAPPED_RESULTS")) == 0);
    c->dChroma = NULL; c->dWp = NULL; c->dWpCap = 0; c->wpReady = false;
    c->dPropagate = NULL; c->dPropOut = NULL; c->dPropOutCap = 0; c->hPropOut = NULL; c->hPropOutCap = 0;
    c->timing = false;
    memset(&c->stats, 0, sizeof(c->stats));
    c->stream = NULL; c->ownStream = false; c->copyStream = NULL; c->evKernel = NULL;
    c->pb = cfg->bitDepth > 8 ? 2 : 1;
    c->pixelMax = (1 << cfg->bitDepth) - 1;
    c->correction = 14 - cfg->bitDepth;
    c->bf = cfg->bframes;
    c->searchWarps = cfg->searchWarps > 0 ? cfg->searchWarps : 4;    /* CU rows (= warps) per commit CTA */
    if (c->searchWarps > SEARCH_MAX_GROUP_ROWS) c->searchWarps = SEARCH_MAX_GROUP_ROWS;
    c->searchSpec = 3;
    c->searchMode = 2; c->specMaxDist = 2; c->specMaxPlans = 4; c->plainWarps = 4;
    if (const char* e = getenv("X265CU_SEARCH_MODE")) c->searchMode = atoi(e);
    if (const char* e = getenv("X265CU_SPEC_MAX_DIST")) c->specMaxDist = atoi(e);
    if (const char* e = getenv("X265CU_SPEC_MAX_PLANS")) c->specMaxPlans = atoi(e);
    c->plainWin = -2;
    if (const char* e = getenv("X265CU_PLAIN_WIN")) c->plainWin = atoi(e);
    if (const char* e = getenv("X265CU_PLAIN_ROWS")) c->plainWarps = atoi(e);
    c->plainTma = 0;
    if (const char* e = getenv("X265CU_PLAIN_TMA")) c->plainTma = atoi(e) != 0;
    c->tmaMapValid = false;
    c->plainOneShot = 0;      /* measured: no gain at 1080p, 10 % slower at 4K (profiles/README.md, round 2) */
    if (const char* e = getenv("X265CU_PLAIN_ONESHOT")) c->plainOneShot = atoi(e) != 0;
    c->plainOct = 0; c->octWarps = 2;
    if (const char* e = getenv("X265CU_PLAIN_OCT")) c->plainOct = atoi(e) != 0;
    if (const char* e = getenv("X265CU_OCT_WARPS")) c->octWarps = atoi(e);
    c->octSlack = 4; c->octSleep = 32; c->octSleepFull = 256; c->octFullWarps = 148 * 8;
    if (const char* e = getenv("X265CU_OCT_SLACK")) c->octSlack = atoi(e);
    if (const char* e = getenv("X265CU_OCT_SLEEP")) c->octSleep = atoi(e);
    if (const char* e = getenv("X265CU_OCT_SLEEP_FULL")) c->octSleepFull = atoi(e);
    if (const char* e = getenv("X265CU_OCT_FULL_WARPS")) c->octFullWarps = atoi(e);
    if (c->octWarps < 1) c->octWarps = 1;
    if (c->octWarps > OCT_MAX_WARPS) c->octWarps = OCT_MAX_WARPS;
    {
        const int smemMax = 96 * 1024;
        cudaFuncSetAttribute(oct_search_kernel<uint8_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax);
        cudaFuncSetAttribute(oct_search_kernel<uint8_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax);
        cudaFuncSetAttribute(oct_search_kernel<uint16_t, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax);
        cudaFuncSetAttribute(oct_search_kernel<uint16_t, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smemMax);
        while (c->octWarps > 1 && oct_smem_bytes<uint16_t>(c->octWarps, (cfg->srcWidth / 2 + 7) / 8) > (size_t)smemMax) c->octWarps--;
    }
    if (c->plainWarps < 1) c->plainWarps = 1;
    if (c->plainWarps > PLAIN_MAX_GROUP_ROWS) c->plainWarps = PLAIN_MAX_GROUP_ROWS;
    memset(c->dbgPlans, 0, sizeof(c->dbgPlans));
    memset(c->hostMs, 0, sizeof(c->hostMs)); memset(c->planMs, 0, sizeof(c->planMs)); c->hostCalls = 0;
    if (const char* e = getenv("X265CU_SEARCH_SPEC")) c->searchSpec = atoi(e);   /* tuning experiments only */
    if (search_smem_bytes<uint16_t>(c->searchWarps, (cfg->srcWidth / 2 + 7) / 8) > 48 * 1024)
    {
        delete c;
        snprintf(g_openError, sizeof(g_openError), "x265cu_open: picture too wide for the search kernel's MV ring");
        return X265CU_EINVAL;
    }

    /* geometry: Lowres::create, common/lowres.cpp:34-48 */
    GeomDev& g = c->g;
    int w = cfg->srcWidth / 2, l = cfg->srcHeight / 2;
    g.stride = w + 2 * cfg->marginX;
    if (g.stride & 31) g.stride += 32 - (g.stride & 31);
    g.wCU = (w + 7) >> 3; g.hCU = (l + 7) >> 3; g.nCU = g.wCU * g.hCU;
    g.width = g.wCU * 8; g.lines = g.hCU * 8;
    g.marginX = cfg->marginX; g.marginY = cfg->marginY;
    g.paddedLines = g.lines + 2 * cfg->marginY;
    g.planeSize = (int64_t)g.stride * g.paddedLines;
    g.padOffset = (int64_t)g.stride * cfg->marginY + cfg->marginX;

#define OPEN_TRY(call)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (call);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
        {                                                                                           \
            snprintf(g_openError, sizeof(g_openError), "x265cu_open: %s: %s", #call, cudaGetErrorString(e_)); \
            freeAll(c); delete c;                                                                   \
            return e_ == cudaErrorMemoryAllocation ? X265CU_ENOMEM : X265CU_ECUDA;                  \
        }                                                                                           \
    } while (0)

    OPEN_TRY(cudaSetDevice(cfg->device));
    g_defaultDevice.store(cfg->device);
    if (cfg->stream) c->stream = (cudaStream_t)cfg->stream;
    else { OPEN_TRY(takeStream(c->cfg.device, &c->stream)); c->ownStream = true; }
    OPEN_TRY(takeStream(c->cfg.device, &c->copyStream));
    OPEN_TRY(takeEvent(c->cfg.device, &c->evKernel));
    c->planesCopied.assign(cfg->numFrameSlots, (cudaEvent_t)NULL);
    c->planesPending.assign(cfg->numFrameSlots, 0);
    c->planesReady.assign(cfg->numFrameSlots, (cudaEvent_t)NULL);
    c->deferPlanes = false;
    {
        x265cu_ctx::SlotUpload z;
        memset(&z, 0, sizeof(z));
        c->slotUp.assign(cfg->numFrameSlots, z);
    }
    for (int i = 0; i < cfg->numFrameSlots; i++)
        OPEN_TRY(takeEvent(c->cfg.device, &c->planesReady[i]));
    for (int i = 0; i < cfg->numFrameSlots; i++)
        OPEN_TRY(takeEvent(c->cfg.device, &c->planesCopied[i]));

    const size_t S = (size_t)cfg->numFrameSlots, n = (size_t)g.nCU, t2 = (size_t)(c->bf + 2) * (c->bf + 2), t1 = (size_t)2 * (c->bf + 1);
    size_t planeBytes = S * 4 * (size_t)g.planeSize * c->pb + 256;
    OPEN_TRY(fixedAlloc(c, (void**)&c->dPlanes, planeBytes));
    {
        /* the frame mirrors as ONE tensor for TMA: dim 0 = samples of a padded row, dim 1 = padded rows of a plane, dim 2 = planes
         * (4 per slot); a search window is the box {WIN_W, WIN_H, 4}.  The driver entry point is looked up at run time (no
         * link-time dependency on libcuda); without it the TMA variant of the plain kernel is simply not offered. */
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                     const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = NULL;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn && qres == cudaDriverEntryPointSuccess)
        {
            const cuuint64_t dims[3] = { (cuuint64_t)c->g.stride, (cuuint64_t)c->g.paddedLines, (cuuint64_t)4 * cfg->numFrameSlots };
            const cuuint64_t strides[2] = { (cuuint64_t)c->g.stride * c->pb, (cuuint64_t)c->g.planeSize * c->pb };
            const cuuint32_t box[3] = { (cuuint32_t)(c->pb == 1 ? PlainTma<uint8_t>::RU * 4 : PlainTma<uint16_t>::RU * 4), WIN_H, 4 };
            const cuuint32_t estr[3] = { 1, 1, 1 };
            CUtensorMap m;
            CUresult r = ((EncodeFn)fn)(&m, c->pb == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_UINT16, 3, c->dPlanes, dims, strides, box, estr,
                                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r == CUDA_SUCCESS) { memcpy(&c->tmaMap, &m, sizeof(m)); c->tmaMapValid = true; }
        }
        else
            cudaGetLastError();
    }
    OPEN_TRY(cudaMemsetAsync(c->dPlanes, 0, planeBytes, c->stream));   /* CHECKED_MALLOC_ZERO, lowres.cpp:62 */
    OPEN_TRY(fixedAlloc(c, (void**)&c->dIntraCost, S * n * sizeof(int)));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dIntraMode, S * n));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dInvQ, S * n * sizeof(int)));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dLowresCosts, S * t2 * n * sizeof(uint16_t)));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dRowSatds, S * t2 * g.hCU * sizeof(int)));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dMvs, S * t1 * n * sizeof(int)));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dMvCosts, S * t1 * n * sizeof(int)));
    OPEN_TRY(cudaMemsetAsync(c->dMvs, 0, S * t1 * n * sizeof(int), c->stream));
    OPEN_TRY(cudaMemsetAsync(c->dMvCosts, 0, S * t1 * n * sizeof(int), c->stream));
    {
        const int bxN = (cfg->srcWidth + 15) / 16, byN = (cfg->srcHeight + 15) / 16;
        c->chromaPitch = 8 * bxN; c->chromaRows = 8 * byN;
        OPEN_TRY(fixedAlloc(c, (void**)&c->dChroma, S * 2 * (size_t)c->chromaPitch * c->chromaRows * c->pb));
        c->hasChroma.assign(S, 0);
    }
    OPEN_TRY(fixedAlloc(c, (void**)&c->dPropagate, S * n * sizeof(unsigned long long)));
    OPEN_TRY(cudaMemsetAsync(c->dPropagate, 0, S * n * sizeof(unsigned long long), c->stream));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dLut, (4 * 32768 + 1) * sizeof(uint16_t)));
    OPEN_TRY(cudaMemcpyAsync(c->dLut, cfg->mvcost - 2 * 32768, (4 * 32768 + 1) * sizeof(uint16_t), cudaMemcpyHostToDevice, c->stream));
    c->srcPitch = (int64_t)alignUp((size_t)(2 * g.width + 1), 64);
    OPEN_TRY(fixedAlloc(c, (void**)&c->dSrc, (size_t)c->srcPitch * (2 * g.lines + 1) * c->pb + 256));
    OPEN_TRY(fixedAlloc(c, (void**)&c->dSmall, 64 * sizeof(unsigned long long)));
    c->hasInvQ.assign(S, 0);
    c->slotPoc.assign(S, 0);
    c->pocBase = 0;
    c->slotPocKnown.assign(S, 0);
    c->mvValid.assign(S * t1, 0);
    OPEN_TRY(cudaStreamSynchronize(c->stream));
    c->cfg.mvcost = NULL;   /* caller's table was copied */
    *out = c;
    return X265CU_OK;
}

void x265cu_close(x265cu_ctx* c)
{
    SlowLog slow("x265cu_close");
    if (!c) return;
    cudaSetDevice(c->cfg.device);
    cudaStreamSynchronize(c->stream);
    if (c->copyStream) cudaStreamSynchronize(c->copyStream);
    if (c->upStream) cudaStreamSynchronize(c->upStream);
    if (c->intraStream) cudaStreamSynchronize(c->intraStream);
    if (getenv("X265CU_HOST_PROFILE"))
        fprintf(stderr, "x265cu_estimate_batch host time over %lld calls: planning %.2f ms (jobs %.2f, hints %.2f, items %.2f, args %.2f), enqueue %.2f ms, wait %.2f ms, scatter %.2f ms\n",
                c->hostCalls, c->hostMs[0], c->planMs[0], c->planMs[1], c->planMs[2], c->planMs[3], c->hostMs[1], c->hostMs[2], c->hostMs[3]);
#ifdef X265CU_PLAIN_CLOCKS
    {
        unsigned long long h[16];
        cudaMemcpyFromSymbol(h, g_plainClk, sizeof(h));
        static const char* nm[10] = { "loop", "fenc+prefetch+neighbours(wait)", "window", "begin+CAND", "START", "HEX", "SQ8", "HPEL", "QPEL", "publish" };
        const double n = h[15] ? (double)h[15] : 1.0;
        fprintf(stderr, "plain kernel, cycles per CU step of warp 0 (%llu steps):", h[15]);
        double tot = 0;
        for (int i = 0; i < 10; i++) { fprintf(stderr, " %s %.0f;", nm[i], h[i] / n); tot += h[i] / n; }
        fprintf(stderr, " total %.0f\n", tot);
        {
            unsigned long long hb[16], hc[8];
            cudaMemcpyFromSymbol(hb, g_plainClkB, sizeof(hb)); cudaMemcpyFromSymbol(hc, g_plainCnt, sizeof(hc));
            const double nb = hb[15] ? (double)hb[15] : 1.0;
            fprintf(stderr, "  the picture's bottom CU row (%llu steps):", hb[15]);
            double tb = 0;
            for (int i = 0; i < 10; i++) { fprintf(stderr, " %s %.0f;", nm[i], hb[i] / nb); tb += hb[i] / nb; }
            fprintf(stderr, " total %.0f | per step: CAND passes %.2f, HEX3 rounds %.2f\n", tb, hc[1] / nb, hc[2] / nb);
        }
        if (const char* path = getenv("X265CU_PLAIN_TRACE"))
        {
            static unsigned long long tt[160][256]; static unsigned int tw[160][256];
            cudaMemcpyFromSymbol(tt, g_plainTraceT, sizeof(tt)); cudaMemcpyFromSymbol(tw, g_plainTraceW, sizeof(tw));
            if (FILE* f = fopen(path, "w"))
            {
                unsigned long long t0 = ~0ull;
                for (int r = 0; r < 160; r++) for (int i = 0; i < 256; i++) if (tt[r][i] && tt[r][i] < t0) t0 = tt[r][i];
                for (int r = 0; r < 160; r++)
                {
                    if (!tt[r][0]) continue;
                    fprintf(f, "row %d:", r);
                    for (int i = 0; i < 256 && tt[r][i]; i++) fprintf(f, " %llu/%u", tt[r][i] - t0, tw[r][i]);
                    fprintf(f, "\n");
                }
                fclose(f);
            }
        }
    }
#endif
#ifdef X265CU_SEARCH_STATS
    {
        unsigned long long h[32];
        cudaMemcpyFromSymbol(h, g_searchStats, sizeof(h));
        fprintf(stderr, "plans %lld hinted(earlier batch) %lld hinted(in-batch seed) %lld\n", c->dbgPlans[0], c->dbgPlans[1], c->dbgPlans[2]);
        fprintf(stderr, "search stats: cus %llu chain searches %llu (%llu, %llu) resume[done,hex6,hex3,sq8,hpel,qpel] %llu %llu %llu %llu %llu %llu\n"
                        "  cycles per CU: total %.0f search %.0f (%.0f per search) wait-below %.0f wait-right %.0f; cand passes %llu fast path %llu\n",
                h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7], h[8], h[9],
                (double)h[13] / (h[0] ? h[0] : 1), (double)h[10] / (h[0] ? h[0] : 1), (double)h[10] / (h[1] ? h[1] : 1),
                (double)h[11] / (h[0] ? h[0] : 1), (double)h[12] / (h[0] ? h[0] : 1), h[14], h[15]);
    }
#endif
    resolveEvents(c);
    freeAll(c);
    delete c;
}

int x265cu_get_geometry(const x265cu_ctx* c, x265cu_geometry* o)
{
    if (!c || !o) return X265CU_EINVAL;
    o->width = c->g.width; o->lines = c->g.lines; o->stride = c->g.stride; o->paddedLines = c->g.paddedLines;
    o->widthInCU = c->g.wCU; o->heightInCU = c->g.hCU; o->cuCount = c->g.nCU;
    o->planeSize = c->g.planeSize; o->padOffset = c->g.padOffset; o->pixelBytes = c->pb;
    return X265CU_OK;
}

int x265cu_sync(x265cu_ctx* c)
{
    if (!c) return X265CU_EINVAL;
    SlowLog slowAll("x265cu_sync");
    std::unique_lock<std::mutex> lk(c->mtx, std::defer_lock);
    { SlowLog slow("x265cu_sync: lock"); lk.lock(); }
    /* (every entry binds the calling thread to the context's device first: a thread that has made no other call yet -- a pool
     * worker running a slicetypeDecide that found everything cached -- would otherwise run on the runtime's default device 0 and
     * create a context there: 0.6-1.4 s per process on an 8-GPU box, measured) */
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    if (flushDeferredPlanes(c)) return X265CU_ECUDA;
    const bool diag = SlowLog::threshold() > 0;
    const cudaError_t q0 = diag ? cudaStreamQuery(c->stream) : cudaSuccess;
    const std::chrono::steady_clock::time_point tq = std::chrono::steady_clock::now();
    int r = syncStream(c);
    if (diag && std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tq).count() >= SlowLog::threshold())
        fprintf(stderr, "x265cu slow call [pid %d]: x265cu_sync found the stream %s before the wait; last entry that enqueued: %s\n", (int)getpid(),
                q0 == cudaSuccess ? "IDLE" : (q0 == cudaErrorNotReady ? "BUSY" : cudaGetErrorString(q0)), c->lastEnqueued ? c->lastEnqueued : "-");
    if (q0 != cudaSuccess) cudaGetLastError();
    {
        SlowLog slow("x265cu_sync: cudaStreamSynchronize(copyStream)");
        CU_TRY(c, cudaStreamSynchronize(c->copyStream));     /* pending plane copy-backs have landed */
    }
    for (size_t i = 0; i < c->planesPending.size(); i++) c->planesPending[i] = 0;
    return r;
}

int x265cu_host_register(void* ptr, size_t bytes)
{
    SlowLog slow("x265cu_host_register");
    if (!ptr || !bytes) return X265CU_EINVAL;
    bindDefaultDevice();
    /* pinned AND mapped: copies to it are asynchronous, and result arrays can be written into it by a kernel */
    cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
    if (e != cudaSuccess) { cudaGetLastError(); return X265CU_ECUDA; }
    void* dev = NULL;
    if (cudaHostGetDevicePointer(&dev, ptr, 0) == cudaSuccess && dev)
    {
        std::lock_guard<std::mutex> lk(g_regMtx);
        HostRange r = { (uintptr_t)ptr, bytes, (uintptr_t)dev };
        g_reg[(uintptr_t)ptr] = r;
    }
    else
        cudaGetLastError();
    return X265CU_OK;
}

int x265cu_host_unregister(void* ptr)
{
    SlowLog slow("x265cu_host_unregister");
    bindDefaultDevice();
    if (!ptr) return X265CU_EINVAL;
    {
        std::lock_guard<std::mutex> lk(g_regMtx);
        g_reg.erase((uintptr_t)ptr);
    }
    cudaError_t e = cudaHostUnregister(ptr);
    if (e != cudaSuccess) { cudaGetLastError(); return X265CU_ECUDA; }
    return X265CU_OK;
}

void x265cu_trim(void)
{
    std::lock_guard<std::mutex> lk(g_poolMtx);
    for (size_t i = 0; i < g_devPool.size(); i++) { cudaSetDevice(g_devPool[i].device); cudaFree(g_devPool[i].p); }
    for (size_t i = 0; i < g_hostPool.size(); i++) cudaFreeHost(g_hostPool[i].p);
    g_devPool.clear();
    g_hostPool.clear();
    std::lock_guard<std::mutex> lr(g_resMtx);
    for (std::map<int, ResPool>::iterator it = g_resPool.begin(); it != g_resPool.end(); ++it)
    {
        cudaSetDevice(it->first);
        for (size_t i = 0; i < it->second.streams.size(); i++) cudaStreamDestroy(it->second.streams[i]);
        for (size_t i = 0; i < it->second.events.size(); i++) cudaEventDestroy(it->second.events[i]);
        for (size_t i = 0; i < it->second.timingEvents.size(); i++) cudaEventDestroy(it->second.timingEvents[i]);
    }
    g_resPool.clear();
}

int x265cu_stats_enable(x265cu_ctx* c, int timing)
{
    if (!c) return X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    syncStream(c);
    c->timing = timing != 0;
    return X265CU_OK;
}

int x265cu_stats_get(x265cu_ctx* c, x265cu_stats* o, int reset)
{
    if (!c || !o) return X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    int r = syncStream(c);
    *o = c->stats;
    if (reset) memset(&c->stats, 0, sizeof(c->stats));
    return r;
}

/* -------------------------------------------------------------------------------------------- */
static int frameInitImpl(x265cu_ctx* c, int slot, const void* luma, intptr_t srcStride, int lumaIsDevice, void* planesOut,
                         const void* u, const void* v, intptr_t cStride, uint32_t* varEnergy, uint64_t* varSums);

int x265cu_frame_init(x265cu_ctx* c, int slot, const void* luma, intptr_t srcStride, int lumaIsDevice, void* planesOut)
{
    return frameInitImpl(c, slot, luma, srcStride, lumaIsDevice, planesOut, NULL, NULL, 0, NULL, NULL);
}

int x265cu_frame_init_var(x265cu_ctx* c, int slot, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride,
                          int planesAreDevice, void* planesOut, uint32_t* energy, uint64_t sums[6])
{
    if (!energy || !sums || ((u == NULL) != (v == NULL))) return c ? fail(c, X265CU_EINVAL, "x265cu_frame_init_var: bad argument") : X265CU_EINVAL;
    return frameInitImpl(c, slot, y, yStride, planesAreDevice, planesOut, u, v, cStride, energy, sums);
}

/* enqueue Lowres::init (and acEnergyCu's integer part when varEnergy != NULL) of one frame on the ctx stream; no
 * synchronisation: varEnergy / varSums (6 x u64) are filled once the stream has drained.  Caller holds the lock. */
struct Uploaded { const uint8_t *y, *u, *v; };   /* picture already on the device with the host's pitches (batched uploads) */

/* results of a batched list live in per-frame device buffers (zeroed / copied back once for the whole list) */
struct BatchOut { unsigned int* dEnergy; unsigned long long* dSums; };

static int frameInitEnqueue(x265cu_ctx* c, int slot, const void* luma, intptr_t srcStride, int lumaIsDevice, void* planesOut,
                            const void* u, const void* v, intptr_t cStride, uint32_t* varEnergy, unsigned long long* varSums,
                            const Uploaded* up = NULL, const BatchOut* bo = NULL)
{
    if (!luma || badSlot(c, slot) || srcStride < 2 * c->g.width + 1) return fail(c, X265CU_EINVAL, "x265cu_frame_init: bad argument");
    const GeomDev& g = c->g;
    const void* src = luma;
    int64_t pitch = srcStride;
    if (up)
    {
        src = up->y;
        pitch = srcStride;
    }
    else if (!lumaIsDevice)
    {
        /* (2W+1) x (2H+1) samples for the downscale; that also covers the 16-aligned picture pixel_var reads */
        size_t wbytes = (size_t)(2 * g.width + 1) * c->pb;
        const size_t rows = (size_t)2 * g.lines + 1;
        if ((((size_t)srcStride * c->pb) & 7) == 0)
        {
            /* ONE linear transfer of the rows with the host's own pitch (the few margin bytes between rows ride along):
             * a strided 2-D copy of ~2 KB rows reaches a fraction of the PCIe rate */
            const size_t lin = (rows - 1) * (size_t)srcStride * c->pb + wbytes;
            if (growDevice(c, &c->dSrcLin, &c->dSrcLinCap, lin + 256)) return X265CU_ECUDA;
            CU_TRY(c, cudaMemcpyAsync(c->dSrcLin, luma, lin, cudaMemcpyHostToDevice, c->stream));
            c->stats.h2dBytes += (int64_t)lin;
            src = c->dSrcLin;
            pitch = srcStride;
        }
        else
        {
            CU_TRY(c, cudaMemcpy2DAsync(c->dSrc, (size_t)c->srcPitch * c->pb, luma, (size_t)srcStride * c->pb, wbytes, rows,
                                        cudaMemcpyHostToDevice, c->stream));
            c->stats.h2dBytes += (int64_t)wbytes * rows;
            src = c->dSrc;
            pitch = c->srcPitch;
        }
    }
    else if (((uintptr_t)luma & 7) || (((size_t)srcStride * c->pb) & 7))
        return fail(c, X265CU_EINVAL, "x265cu_frame_init: device luma must be 8-byte aligned with an 8-byte multiple pitch");
    for (size_t k = 0; k < c->deferredPlanes.size(); k++)
        if (c->deferredPlanes[k].slot == slot) { int fr = flushDeferredPlanes(c); if (fr) return fr; break; }   /* a held-back copy of this slot's old planes */
    if (c->planesPending[slot])
        CU_TRY(c, cudaStreamWaitEvent(c->stream, c->planesCopied[slot], 0));   /* do not overwrite planes still being copied out */
    /* a new picture in this slot: its order and its MV fields are unknown again */
    c->slotPocKnown[slot] = 0;
    c->hasChroma[slot] = 0;             /* set again below when this call measures the variance with chroma planes */
    std::fill(c->mvValid.begin() + (size_t)slot * 2 * (c->bf + 1), c->mvValid.begin() + (size_t)(slot + 1) * 2 * (c->bf + 1), 0);
    {
        KernelScope ks(c, X265CU_K_LOWRES);
        const int padW = g.width + 2 * g.marginX, unit = 16 / c->pb;
        dim3 grid((unsigned)(((int64_t)((padW + unit - 1) / unit) * g.paddedLines + 255) / 256));
        if (c->pb == 1)
            lowres_init_kernel<uint8_t><<<grid, 256, 0, c->stream>>>((const uint8_t*)src, pitch, (uint8_t*)slotBuffer(c, slot), g);
        else
            lowres_init_kernel<uint16_t><<<grid, 256, 0, c->stream>>>((const uint16_t*)src, pitch, (uint16_t*)slotBuffer(c, slot), g);
    }
    CU_TRY(c, cudaGetLastError());
    if (planesOut && c->deferPlanes)
    {
        CU_TRY(c, cudaEventRecord(c->planesReady[slot], c->stream));
        x265cu_ctx::DeferredPlanes d = { slot, planesOut };
        c->deferredPlanes.push_back(d);
    }
    else if (planesOut)
    {
        /* the planes travel back on the copy stream, behind the compute stream: complete after x265cu_sync() */
        size_t bytes = (size_t)4 * g.planeSize * c->pb;
        CU_TRY(c, cudaEventRecord(c->evKernel, c->stream));
        CU_TRY(c, cudaStreamWaitEvent(c->copyStream, c->evKernel, 0));
        CU_TRY(c, cudaMemcpyAsync(planesOut, slotBuffer(c, slot), bytes, cudaMemcpyDeviceToHost, c->copyStream));
        CU_TRY(c, cudaEventRecord(c->planesCopied[slot], c->copyStream));
        c->planesPending[slot] = 1;
        c->stats.d2hBytes += (int64_t)bytes;
    }
    if (varEnergy)
    {
        /* acEnergyCu's integer work on the luma that is already on the device + the two chroma planes */
        const int W = c->cfg.srcWidth, H = c->cfg.srcHeight;
        const int bxN = (W + 15) / 16, byN = (H + 15) / 16;
        /* chroma rows travel with the host's pitch in one linear transfer per plane, like the luma */
        const size_t cp = (u && !lumaIsDevice) ? (size_t)cStride : alignUp((size_t)bxN * 8, 64);
        const size_t cLin = ((size_t)byN * 8 - 1) * cp * c->pb + (size_t)bxN * 8 * c->pb;
        const size_t cBytes = alignUp(cLin, 256);
        const size_t eBytes = alignUp((size_t)bxN * byN * 4, 256);
        if (growDevice(c, &c->dGeneric, &c->dGenericCap, 2 * cBytes + eBytes + 256)) return X265CU_ECUDA;
        uint8_t* dU = c->dGeneric; uint8_t* dV = dU + cBytes; unsigned int* dE = (unsigned int*)(dV + cBytes);
        int64_t cpitch = (int64_t)cp;
        if (u && up) { dU = (uint8_t*)up->u; dV = (uint8_t*)up->v; }
        else if (u && !lumaIsDevice)
        {
            CU_TRY(c, cudaMemcpyAsync(dU, u, cLin, cudaMemcpyHostToDevice, c->stream));
            CU_TRY(c, cudaMemcpyAsync(dV, v, cLin, cudaMemcpyHostToDevice, c->stream));
            c->stats.h2dBytes += (int64_t)(2 * cLin);
        }
        else if (u) { dU = (uint8_t*)u; dV = (uint8_t*)v; cpitch = cStride; }
        unsigned long long* dSums = c->dSmall;
        if (bo) { dE = bo->dEnergy; dSums = bo->dSums; }
        else CU_TRY(c, cudaMemsetAsync(c->dSmall, 0, 6 * sizeof(unsigned long long), c->stream));
        {
            KernelScope ks(c, X265CU_K_VAR);
            int blocks = (bxN * byN + 7) / 8;
            if (blocks > 148 * 8) blocks = 148 * 8;      /* warps stride over the 16x16 blocks */
            if (c->pb == 1)
                frame_var_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>((const uint8_t*)src, pitch, u ? (const uint8_t*)dU : NULL, u ? (const uint8_t*)dV : NULL, cpitch, bxN, byN, dE, dSums,
                                                                          u ? (uint8_t*)slotChroma(c, slot, 0) : NULL, u ? (uint8_t*)slotChroma(c, slot, 1) : NULL);
            else
                frame_var_kernel<uint16_t><<<blocks, 256, 0, c->stream>>>((const uint16_t*)src, pitch, u ? (const uint16_t*)dU : NULL, u ? (const uint16_t*)dV : NULL, cpitch, bxN, byN, dE, dSums,
                                                                           u ? (uint16_t*)slotChroma(c, slot, 0) : NULL, u ? (uint16_t*)slotChroma(c, slot, 1) : NULL);
            c->hasChroma[slot] = u != NULL;
        }
        CU_TRY(c, cudaGetLastError());
        if (!bo)
        {
            CU_TRY(c, cudaMemcpyAsync(varEnergy, dE, (size_t)bxN * byN * 4, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(c, cudaMemcpyAsync(varSums, c->dSmall, 6 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        }
    }
    return X265CU_OK;
}

static int frameInitImpl(x265cu_ctx* c, int slot, const void* luma, intptr_t srcStride, int lumaIsDevice, void* planesOut,
                         const void* u, const void* v, intptr_t cStride, uint32_t* varEnergy, uint64_t* varSums)
{
    if (!c) return X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    unsigned long long hs[6] = { 0, 0, 0, 0, 0, 0 };
    if (!badSlot(c, slot)) c->slotUp[slot].valid = false;      /* this initialisation supersedes a picture uploaded ahead */
    int r = frameInitEnqueue(c, slot, luma, srcStride, lumaIsDevice, planesOut, u, v, cStride, varEnergy, hs);
    if (r) return r;
    /* the caller may reuse its picture buffers as soon as we return */
    r = syncStream(c);
    if (varEnergy) for (int i = 0; i < 6; i++) varSums[i] = hs[i];
    return r;
}

/* PreLookaheadGroup::processTasks hands a LIST of frames to the workers (slicetype.cpp:831-856): all uploads and
 * kernels of the list are enqueued back to back and the host waits once */
int x265cu_frame_upload(x265cu_ctx* c, int slot, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride)
{
    if (!c || badSlot(c, slot) || !y || !u || !v) return c ? fail(c, X265CU_EINVAL, "x265cu_frame_upload: bad argument") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const GeomDev& g = c->g;
    if ((((size_t)yStride * c->pb) & 7) || yStride < 2 * g.width + 1)
        return X265CU_OK;                       /* a pitch the linear transfer cannot take: the pre-lookahead uploads as usual */
    const int bxN = (c->cfg.srcWidth + 15) / 16, byN = (c->cfg.srcHeight + 15) / 16;
    const size_t lumaRows = (size_t)2 * g.lines + 1, chromaRows = (size_t)byN * 8;
    const size_t yLin = (lumaRows - 1) * (size_t)yStride * c->pb + (size_t)(2 * g.width + 1) * c->pb;
    const size_t cLin = (chromaRows - 1) * (size_t)cStride * c->pb + (size_t)bxN * 8 * c->pb;
    const size_t need = alignUp(yLin, 256) + 2 * alignUp(cLin, 256) + 256;
    x265cu_ctx::SlotUpload& su = c->slotUp[slot];
    if (!c->upStream) CU_TRY(c, takeStream(c->cfg.device, &c->upStream));
    if (!su.done) { CU_TRY(c, takeEvent(c->cfg.device, &su.done)); CU_TRY(c, takeEvent(c->cfg.device, &su.read)); }
    if (su.cap < need)
    {
        if (su.d) { CU_TRY(c, cudaStreamSynchronize(c->stream)); CU_TRY(c, cudaStreamSynchronize(c->upStream)); poolGive(g_devPool, c->cfg.device, su.d, su.cap); su.d = NULL; su.cap = 0; }
        size_t got = 0;
        void* sp = poolTake(g_devPool, c->cfg.device, need, &got);
        if (sp && got > 2 * need) { poolGive(g_devPool, c->cfg.device, sp, got); sp = NULL; }
        if (sp) { su.d = (uint8_t*)sp; su.cap = got; }
        else { CU_TRY(c, cudaMalloc((void**)&su.d, need)); su.cap = need; }
        su.everRead = false;
    }
    /* the kernels that read the slot's previous picture from this area must be done with it */
    if (su.everRead) CU_TRY(c, cudaStreamWaitEvent(c->upStream, su.read, 0));
    uint8_t* dY = su.d; uint8_t* dU = dY + alignUp(yLin, 256); uint8_t* dV = dU + alignUp(cLin, 256);
    CU_TRY(c, cudaMemcpyAsync(dY, y, yLin, cudaMemcpyHostToDevice, c->upStream));
    CU_TRY(c, cudaMemcpyAsync(dU, u, cLin, cudaMemcpyHostToDevice, c->upStream));
    CU_TRY(c, cudaMemcpyAsync(dV, v, cLin, cudaMemcpyHostToDevice, c->upStream));
    CU_TRY(c, cudaEventRecord(su.done, c->upStream));
    c->stats.h2dBytes += (int64_t)(yLin + 2 * cLin);
    su.y = y; su.u = u; su.v = v; su.ys = yStride; su.cs = cStride; su.valid = true;
    return X265CU_OK;
}

static int preBatchImpl(x265cu_ctx* c, int n, const x265cu_frame_in* items, x265cu_aq_fn aq, void* user, x265cu_intra_out* outs);
static void weightArgs(const x265cu_ctx* c, int scale, int denom, int offset, int* round, int* shift, int* off);
static int intraEnqueue(x265cu_ctx* c, int slot, x265cu_intra_out* out, unsigned long long* sums, unsigned long long* dBatchSums, cudaStream_t st);
static int intraEnqueueBatch(x265cu_ctx* c, int count, const int* slots, x265cu_intra_out* const* outs, unsigned long long* dBatchSums, cudaStream_t st);

int x265cu_frame_init_var_batch(x265cu_ctx* c, int n, const x265cu_frame_in* items)
{
    if (!c || n < 0 || (n && !items)) return c ? fail(c, X265CU_EINVAL, "x265cu_frame_init_var_batch: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    return preBatchImpl(c, n, items, NULL, NULL, NULL);
}

int x265cu_pre_lookahead_batch(x265cu_ctx* c, int n, const x265cu_frame_in* items, x265cu_aq_fn aq, void* user, x265cu_intra_out* outs)
{
    if (c) c->lastEnqueued = "pre_lookahead_batch";
    SlowLog slow("x265cu_pre_lookahead_batch");
    if (!c || n < 0 || (n && (!items || !aq || !outs))) return c ? fail(c, X265CU_EINVAL, "x265cu_pre_lookahead_batch: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    return preBatchImpl(c, n, items, aq, user, outs);
}

/* aq == NULL: lowres planes + variance of the list, one wait.  aq != NULL: per frame, as soon as its energies/sums
 * are on the host, the caller's float AQ mapping, then invQscaleFactor upload + intra estimate on the intra stream */
static int preBatchImpl2(x265cu_ctx* c, int n, const x265cu_frame_in* items, x265cu_aq_fn aq, void* user, x265cu_intra_out* outs);
static int preBatchImpl(x265cu_ctx* c, int n, const x265cu_frame_in* items, x265cu_aq_fn aq, void* user, x265cu_intra_out* outs)
{
    c->deferPlanes = !(getenv("X265CU_DEFER_PLANES") && atoi(getenv("X265CU_DEFER_PLANES")) == 0);
    int r = preBatchImpl2(c, n, items, aq, user, outs);
    c->deferPlanes = false;
    return r;
}
static int preBatchImpl2(x265cu_ctx* c, int n, const x265cu_frame_in* items, x265cu_aq_fn aq, void* user, x265cu_intra_out* outs)
{
    const int bxN = (c->cfg.srcWidth + 15) / 16, byN = (c->cfg.srcHeight + 15) / 16;
    const size_t eBytes = alignUp((size_t)bxN * byN * 4, 64), per = eBytes + 64;
    const size_t intraSumsOff = alignUp((size_t)n * per, 256);      /* [n x 16 bytes] intra sums behind the var records */
    if (growHost(c, &c->hPre, &c->hPreCap, intraSumsOff + (size_t)n * 16)) return X265CU_ECUDA;
    /* device mirror of that landing zone: [frame][energy | 6 sums]; zeroed and copied back once for the whole list */
    if (growDevice(c, &c->dPre, &c->dPreCap, intraSumsOff + (size_t)n * 16)) return X265CU_ECUDA;
    CU_TRY(c, cudaMemsetAsync(c->dPre, 0, intraSumsOff + (size_t)n * 16, c->stream));
    if (aq)
    {
        if (!c->intraStream) CU_TRY(c, takeStream(c->cfg.device, &c->intraStream));
        while ((int)c->preEvents.size() < n)
        {
            cudaEvent_t e;
            CU_TRY(c, takeEvent(c->cfg.device, &e));
            c->preEvents.push_back(e);
        }
    }
    const bool dbgPre = getenv("X265CU_PRE_DEBUG") != NULL;
    std::chrono::steady_clock::time_point tA = std::chrono::steady_clock::now(), tB = tA, tC = tA;
    /* host pictures: every frame gets its own staging area and its uploads run on the upload stream, so frame i + 1
     * crosses PCIe while frame i is being downscaled / measured (one linear transfer per plane, host pitch kept) */
    const GeomDev& g = c->g;
    const size_t lumaRows = (size_t)2 * g.lines + 1, chromaRows = (size_t)byN * 8;
    bool pipelined = true;
    std::vector<size_t> off((size_t)n + 1, 0);
    for (int i = 0; i < n; i++)
    {
        const x265cu_frame_in& f = items[i];
        if (!f.energy || !f.sums || ((f.u == NULL) != (f.v == NULL)) || !f.y || badSlot(c, f.slot)) return fail(c, X265CU_EINVAL, "x265cu_frame_init_var_batch: bad item");
        if (f.planesAreDevice || !f.u || (((size_t)f.yStride * c->pb) & 7) || f.yStride < 2 * g.width + 1) pipelined = false;
        const size_t yLin = (lumaRows - 1) * (size_t)f.yStride * c->pb + (size_t)(2 * g.width + 1) * c->pb;
        const size_t cLin = f.u ? (chromaRows - 1) * (size_t)f.cStride * c->pb + (size_t)bxN * 8 * c->pb : 0;
        off[i + 1] = off[i] + alignUp(yLin, 256) + 2 * alignUp(cLin, 256);
    }
    if (pipelined)
    {
        if (growDevice(c, &c->dUp, &c->dUpCap, off[n] + 256)) return X265CU_ECUDA;
        if (!c->upStream) CU_TRY(c, takeStream(c->cfg.device, &c->upStream));
        while ((int)c->upEvents.size() < n)
        {
            cudaEvent_t e;
            CU_TRY(c, takeEvent(c->cfg.device, &e));
            c->upEvents.push_back(e);
        }
        for (int i = 0; i < n; i++)
        {
            const x265cu_frame_in& f = items[i];
            const size_t yLin = (lumaRows - 1) * (size_t)f.yStride * c->pb + (size_t)(2 * g.width + 1) * c->pb;
            const size_t cLin = (chromaRows - 1) * (size_t)f.cStride * c->pb + (size_t)bxN * 8 * c->pb;
            const x265cu_ctx::SlotUpload& su = c->slotUp[f.slot];
            if (su.valid && su.y == f.y && su.u == f.u && su.v == f.v && su.ys == f.yStride && su.cs == f.cStride)
                continue;                       /* uploaded ahead by x265cu_frame_upload */
            uint8_t* dY = c->dUp + off[i]; uint8_t* dU = dY + alignUp(yLin, 256); uint8_t* dV = dU + alignUp(cLin, 256);
            CU_TRY(c, cudaMemcpyAsync(dY, f.y, yLin, cudaMemcpyHostToDevice, c->upStream));
            CU_TRY(c, cudaMemcpyAsync(dU, f.u, cLin, cudaMemcpyHostToDevice, c->upStream));
            CU_TRY(c, cudaMemcpyAsync(dV, f.v, cLin, cudaMemcpyHostToDevice, c->upStream));
            CU_TRY(c, cudaEventRecord(c->upEvents[i], c->upStream));
            c->stats.h2dBytes += (int64_t)(yLin + 2 * cLin);
        }
    }
    tB = std::chrono::steady_clock::now();
    /* Frames whose picture is on the device (uploaded through the staging areas, or device pointers) go through the kernels
     * a CHUNK at a time: one lowres launch (blockIdx.z = frame) and one variance launch (blockIdx.y = frame) per chunk, one
     * copy of the chunk's records, one event.  Host mode keeps chunks short so that the first kernels start while the later
     * pictures are still crossing PCIe. */
    bool chunked = pipelined;
    if (!pipelined)
    {
        chunked = true;
        for (int i = 0; i < n; i++)
            if (!items[i].planesAreDevice || ((uintptr_t)items[i].y & 7) || (((size_t)items[i].yStride * c->pb) & 7)) chunked = false;
    }
    if (getenv("X265CU_PRE_CHUNKED") && atoi(getenv("X265CU_PRE_CHUNKED")) == 0) chunked = false;
    const int chunkMax = !chunked ? 1 : (pipelined ? (n >= 12 ? 4 : 2) : PRE_BATCH);
    std::vector<int> chunkFirst;
    for (int first = 0; first < n; first += chunkMax) chunkFirst.push_back(first);
    chunkFirst.push_back(n);
    if (aq)
        while (c->preEvents.size() < chunkFirst.size())
        {
            cudaEvent_t e;
            CU_TRY(c, takeEvent(c->cfg.device, &e));
            c->preEvents.push_back(e);
        }
    for (size_t ch = 0; ch + 1 < chunkFirst.size() && chunked; ch++)
    {
        const int first = chunkFirst[ch], count = chunkFirst[ch + 1] - first;
        LowresBatch lb; VarBatch vb;
        memset(&lb, 0, sizeof(lb)); memset(&vb, 0, sizeof(vb));
        bool staged[PRE_BATCH];
        for (int k2 = 0; k2 < count; k2++)
        {
            const int i = first + k2;
            const x265cu_frame_in& f = items[i];
            if (badSlot(c, f.slot) || f.yStride < 2 * g.width + 1) return fail(c, X265CU_EINVAL, "x265cu_frame_init_var_batch: bad slot or stride");
            const size_t yLin = (lumaRows - 1) * (size_t)f.yStride * c->pb + (size_t)(2 * g.width + 1) * c->pb;
            const size_t cLin = f.u ? (chromaRows - 1) * (size_t)f.cStride * c->pb + (size_t)bxN * 8 * c->pb : 0;
            staged[k2] = false;
            if (pipelined)
            {
                x265cu_ctx::SlotUpload& su = c->slotUp[f.slot];
                staged[k2] = su.valid && su.y == f.y && su.u == f.u && su.v == f.v && su.ys == f.yStride && su.cs == f.cStride;
                const uint8_t* base = staged[k2] ? su.d : c->dUp + off[i];
                lb.src[k2] = base; vb.y[k2] = base; vb.u[k2] = base + alignUp(yLin, 256); vb.v[k2] = base + alignUp(yLin, 256) + alignUp(cLin, 256);
                CU_TRY(c, cudaStreamWaitEvent(c->stream, staged[k2] ? su.done : c->upEvents[i], 0));
            }
            else
            {
                lb.src[k2] = f.y; vb.y[k2] = f.y; vb.u[k2] = f.u; vb.v[k2] = f.v;
            }
            lb.planes[k2] = slotBuffer(c, f.slot);
            lb.pitch[k2] = f.yStride; vb.ys[k2] = f.yStride; vb.cs[k2] = f.cStride;
            vb.uKeep[k2] = f.u ? slotChroma(c, f.slot, 0) : NULL; vb.vKeep[k2] = f.u ? slotChroma(c, f.slot, 1) : NULL;
            c->hasChroma[f.slot] = f.u != NULL;
            vb.energy[k2] = (unsigned int*)(c->dPre + (size_t)i * per);
            vb.sums[k2] = (unsigned long long*)(c->dPre + (size_t)i * per + eBytes);
            for (size_t d = 0; d < c->deferredPlanes.size(); d++)
                if (c->deferredPlanes[d].slot == f.slot) { int fr = flushDeferredPlanes(c); if (fr) return fr; break; }   /* a held-back copy of this slot's old planes */
            if (c->planesPending[f.slot])
                CU_TRY(c, cudaStreamWaitEvent(c->stream, c->planesCopied[f.slot], 0));   /* do not overwrite planes still being copied out */
            c->slotPocKnown[f.slot] = 0;       /* a new picture in this slot: its order and its MV fields are unknown again */
            std::fill(c->mvValid.begin() + (size_t)f.slot * 2 * (c->bf + 1), c->mvValid.begin() + (size_t)(f.slot + 1) * 2 * (c->bf + 1), 0);
        }
        {
            KernelScope ks(c, X265CU_K_LOWRES);
            const int padW = g.width + 2 * g.marginX, unit = 16 / c->pb;
            dim3 grid((unsigned)(((int64_t)((padW + unit - 1) / unit) * g.paddedLines + 255) / 256), 1, count);
            if (c->pb == 1) lowres_init_batch_kernel<uint8_t><<<grid, 256, 0, c->stream>>>(lb, g);
            else lowres_init_batch_kernel<uint16_t><<<grid, 256, 0, c->stream>>>(lb, g);
        }
        CU_TRY(c, cudaGetLastError());
        for (int k2 = 0; k2 < count; k2++)
        {
            const x265cu_frame_in& f = items[first + k2];
            if (f.planesOut && c->deferPlanes)
            {
                CU_TRY(c, cudaEventRecord(c->planesReady[f.slot], c->stream));
                x265cu_ctx::DeferredPlanes d = { f.slot, f.planesOut };
                c->deferredPlanes.push_back(d);
            }
            else if (f.planesOut)
            {
                const size_t bytes = (size_t)4 * g.planeSize * c->pb;
                CU_TRY(c, cudaEventRecord(c->evKernel, c->stream));
                CU_TRY(c, cudaStreamWaitEvent(c->copyStream, c->evKernel, 0));
                CU_TRY(c, cudaMemcpyAsync(f.planesOut, slotBuffer(c, f.slot), bytes, cudaMemcpyDeviceToHost, c->copyStream));
                CU_TRY(c, cudaEventRecord(c->planesCopied[f.slot], c->copyStream));
                c->planesPending[f.slot] = 1;
                c->stats.d2hBytes += (int64_t)bytes;
            }
        }
        {
            KernelScope ks(c, X265CU_K_VAR);
            int blocks = (bxN * byN + 7) / 8;
            const int varPair = getenv("X265CU_VAR_PAIR") ? atoi(getenv("X265CU_VAR_PAIR")) : 1;           /* 0: a quad per block (tests) */
            const int varCtas = getenv("X265CU_VAR_CTAS") && atoi(getenv("X265CU_VAR_CTAS")) >= 148 ? atoi(getenv("X265CU_VAR_CTAS")) : 148 * 8;
            const int cap = varCtas / count > 148 ? varCtas / count : 148;      /* warps stride over the 16x16 blocks */
            if (blocks > cap) blocks = cap;
            dim3 grid(blocks, count);
            if (c->pb == 1 && varPair) frame_var_batch_kernel<uint8_t, true><<<grid, 256, 0, c->stream>>>(vb, bxN, byN);
            else if (c->pb == 1) frame_var_batch_kernel<uint8_t, false><<<grid, 256, 0, c->stream>>>(vb, bxN, byN);
            else if (varPair) frame_var_batch_kernel<uint16_t, true><<<grid, 256, 0, c->stream>>>(vb, bxN, byN);
            else frame_var_batch_kernel<uint16_t, false><<<grid, 256, 0, c->stream>>>(vb, bxN, byN);
        }
        CU_TRY(c, cudaGetLastError());
        for (int k2 = 0; k2 < count; k2++)
        {
            const x265cu_frame_in& f = items[first + k2];
            c->slotUp[f.slot].valid = false;        /* consumed (or superseded by this initialisation) */
            if (staged[k2])
            {
                CU_TRY(c, cudaEventRecord(c->slotUp[f.slot].read, c->stream));
                c->slotUp[f.slot].everRead = true;
            }
        }
        if (aq)
        {
            /* the chunk's records come back in one copy, and an event tells the host (and the intra stream) they are there */
            CU_TRY(c, cudaMemcpyAsync(c->hPre + (size_t)first * per, c->dPre + (size_t)first * per, (size_t)count * per, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(c, cudaEventRecord(c->preEvents[ch], c->stream));
        }
    }
    for (int i = 0; i < n && !chunked; i++)
    {
        const x265cu_frame_in& f = items[i];
        Uploaded up;
        bool staged = false;
        if (pipelined)
        {
            const size_t yLin = (lumaRows - 1) * (size_t)f.yStride * c->pb + (size_t)(2 * g.width + 1) * c->pb;
            const size_t cLin = (chromaRows - 1) * (size_t)f.cStride * c->pb + (size_t)bxN * 8 * c->pb;
            x265cu_ctx::SlotUpload& su = c->slotUp[f.slot];
            staged = su.valid && su.y == f.y && su.u == f.u && su.v == f.v && su.ys == f.yStride && su.cs == f.cStride;
            if (staged)
            {
                up.y = su.d; up.u = up.y + alignUp(yLin, 256); up.v = up.u + alignUp(cLin, 256);
                CU_TRY(c, cudaStreamWaitEvent(c->stream, su.done, 0));
            }
            else
            {
                up.y = c->dUp + off[i]; up.u = up.y + alignUp(yLin, 256); up.v = up.u + alignUp(cLin, 256);
                CU_TRY(c, cudaStreamWaitEvent(c->stream, c->upEvents[i], 0));
            }
        }
        BatchOut bo = { (unsigned int*)(c->dPre + (size_t)i * per), (unsigned long long*)(c->dPre + (size_t)i * per + eBytes) };
        int r = frameInitEnqueue(c, f.slot, f.y, f.yStride, f.planesAreDevice, f.planesOut, f.u, f.v, f.cStride,
                                 (uint32_t*)(c->hPre + (size_t)i * per), (unsigned long long*)(c->hPre + (size_t)i * per + eBytes),
                                 pipelined ? &up : NULL, &bo);
        if (r) { cudaStreamSynchronize(c->stream); if (c->upStream) cudaStreamSynchronize(c->upStream); return r; }
        c->slotUp[f.slot].valid = false;        /* consumed (or superseded by this initialisation) */
        if (staged)
        {
            CU_TRY(c, cudaEventRecord(c->slotUp[f.slot].read, c->stream));
            c->slotUp[f.slot].everRead = true;
        }
        if (aq)
        {
            /* this frame's record comes back on its own, and an event tells the host (and the intra stream) it is done */
            CU_TRY(c, cudaMemcpyAsync(c->hPre + (size_t)i * per, c->dPre + (size_t)i * per, per, cudaMemcpyDeviceToHost, c->stream));
            CU_TRY(c, cudaEventRecord(c->preEvents[i], c->stream));
        }
    }
    if (aq)
    {
        tC = std::chrono::steady_clock::now();
        int rc = X265CU_OK;
        for (size_t ch = 0; ch + 1 < chunkFirst.size() && !rc; ch++)
        {
            const int first = chunkFirst[ch], count = chunkFirst[ch + 1] - first;
            if (cudaEventSynchronize(c->preEvents[chunked ? ch : (size_t)(first + count - 1)]) != cudaSuccess) { rc = fail(c, X265CU_ECUDA, "x265cu_pre_lookahead_batch: event wait failed"); break; }
            const int32_t* invQs[8] = { NULL, NULL, NULL, NULL, NULL, NULL, NULL, NULL };
            for (int i = first; i < first + count; i++)
            {
                memcpy(items[i].energy, c->hPre + (size_t)i * per, (size_t)bxN * byN * 4);
                memcpy(items[i].sums, c->hPre + (size_t)i * per + eBytes, 6 * sizeof(uint64_t));
            }
            aq(user, first, count, invQs);                      /* the host's float AQ mapping of these frames */
            cudaStreamWaitEvent(c->intraStream, c->preEvents[chunked ? ch : (size_t)(first + count - 1)], 0);
            int slots[PRE_BATCH];
            x265cu_intra_out* po[PRE_BATCH];
            for (int i = first; i < first + count && !rc; i++)
            {
                const int32_t* invQ = invQs[i - first];
                const int slot = items[i].slot;
                slots[i - first] = slot; po[i - first] = &outs[i];
                c->hasInvQ[slot] = invQ != NULL;
                if (invQ)
                {
                    if (cudaMemcpyAsync(slotInvQ(c, slot), invQ, (size_t)g.nCU * sizeof(int), cudaMemcpyHostToDevice, c->intraStream) != cudaSuccess)
                    { rc = fail(c, X265CU_ECUDA, "x265cu_pre_lookahead_batch: invQscale upload failed"); break; }
                    c->stats.h2dBytes += (int64_t)g.nCU * sizeof(int);
                }
            }
            if (!rc) rc = intraEnqueueBatch(c, count, slots, po, (unsigned long long*)(c->dPre + intraSumsOff + (size_t)first * 16), c->intraStream);
        }
        if (!rc && cudaMemcpyAsync(c->hPre + intraSumsOff, c->dPre + intraSumsOff, (size_t)n * 16, cudaMemcpyDeviceToHost, c->intraStream) != cudaSuccess)
            rc = fail(c, X265CU_ECUDA, "x265cu_pre_lookahead_batch: copy failed");
        cudaError_t e1 = cudaStreamSynchronize(c->intraStream);
        int r2 = syncStream(c);
        if (rc) return rc;
        if (e1 != cudaSuccess) return fail(c, X265CU_ECUDA, cudaGetErrorString(e1));
        if (r2) return r2;
        for (int i = 0; i < n; i++)
        {
            const unsigned long long* sm = (const unsigned long long*)(c->hPre + intraSumsOff + (size_t)i * 16);
            outs[i].sums[0] = (int64_t)sm[0];
            outs[i].sums[1] = (int64_t)sm[1];
        }
        if (dbgPre)
            fprintf(stderr, "  pre_lookahead_batch n=%d pipelined=%d: upload submit %.2f ms, enqueue %.2f ms, per-frame AQ + intra + wait %.2f ms\n", n, (int)pipelined,
                    std::chrono::duration<double, std::milli>(tB - tA).count(), std::chrono::duration<double, std::milli>(tC - tB).count(),
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tC).count());
        return X265CU_OK;
    }
    CU_TRY(c, cudaMemcpyAsync(c->hPre, c->dPre, (size_t)n * per, cudaMemcpyDeviceToHost, c->stream));
    tC = std::chrono::steady_clock::now();
    int r = syncStream(c);
    if (r) return r;
    if (dbgPre)
        fprintf(stderr, "  frame_init_var_batch n=%d pipelined=%d: upload submit %.2f ms, enqueue %.2f ms, wait %.2f ms\n", n, (int)pipelined,
                std::chrono::duration<double, std::milli>(tB - tA).count(), std::chrono::duration<double, std::milli>(tC - tB).count(),
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tC).count());
    for (int i = 0; i < n; i++)
    {
        memcpy(items[i].energy, c->hPre + (size_t)i * per, (size_t)bxN * byN * 4);
        memcpy(items[i].sums, c->hPre + (size_t)i * per + eBytes, 6 * sizeof(uint64_t));
    }
    return X265CU_OK;
}

int x265cu_frame_set_invqscale(x265cu_ctx* c, int slot, const int32_t* invQ)
{
    if (c) c->lastEnqueued = "frame_set_invqscale";
    if (!c || badSlot(c, slot)) return c ? fail(c, X265CU_EINVAL, "x265cu_frame_set_invqscale: bad slot") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    c->hasInvQ[slot] = invQ != NULL;
    if (invQ)
    {
        /* stream-ordered before any kernel that reads it; Lowres::invQscaleFactor lives as long as the frame */
        CU_TRY(c, cudaMemcpyAsync(slotInvQ(c, slot), invQ, (size_t)c->g.nCU * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        c->stats.h2dBytes += (int64_t)c->g.nCU * sizeof(int);
        return X265CU_OK;
    }
    return X265CU_OK;
}

/* enqueue lowresIntraEstimate of one frame; sums (2 x u64) land once the stream has drained.  Caller holds the lock. */
static int intraEnqueue(x265cu_ctx* c, int slot, x265cu_intra_out* out, unsigned long long* sums, unsigned long long* dBatchSums,
                        cudaStream_t st)
{
    if (!st) st = c->stream;
    if (badSlot(c, slot)) return fail(c, X265CU_EINVAL, "x265cu_intra: bad slot");
    const GeomDev& g = c->g;
    IntraOutDev o;
    o.intraCost = slotIntraCost(c, slot);
    o.intraMode = slotIntraMode(c, slot);
    o.lowresCosts = slotLowresCosts(c, slot, 0, 0);
    o.rowSatds = slotRowSatds(c, slot, 0, 0);
    o.sums = dBatchSums ? dBatchSums : c->dSmall;
    o.invQ = c->hasInvQ[slot] ? slotInvQ(c, slot) : NULL;
    CU_TRY(c, cudaMemsetAsync(o.rowSatds, 0, (size_t)g.hCU * sizeof(int), st));
    if (!dBatchSums) CU_TRY(c, cudaMemsetAsync(c->dSmall, 0, 2 * sizeof(unsigned long long), st));
    {
        KernelScope ks(c, X265CU_K_INTRA, 1, st);
        int blocks = (g.nCU + 7) / 8;
        if (c->pb == 1)
            intra_kernel<uint8_t><<<blocks, 256, 0, st>>>((const uint8_t*)slotPlane0(c, slot), g, c->cfg.lookaheadLambda, c->pixelMax, o);
        else
            intra_kernel<uint16_t><<<blocks, 256, 0, st>>>((const uint16_t*)slotPlane0(c, slot), g, c->cfg.lookaheadLambda, c->pixelMax, o);
    }
    CU_TRY(c, cudaGetLastError());
    if (out)
    {
        if (out->intraCost) { CU_TRY(c, cudaMemcpyAsync(out->intraCost, o.intraCost, (size_t)g.nCU * 4, cudaMemcpyDeviceToHost, st)); c->stats.d2hBytes += g.nCU * 4; }
        if (out->intraMode) { CU_TRY(c, cudaMemcpyAsync(out->intraMode, o.intraMode, (size_t)g.nCU, cudaMemcpyDeviceToHost, st)); c->stats.d2hBytes += g.nCU; }
        if (out->lowresCosts) { CU_TRY(c, cudaMemcpyAsync(out->lowresCosts, o.lowresCosts, (size_t)g.nCU * 2, cudaMemcpyDeviceToHost, st)); c->stats.d2hBytes += g.nCU * 2; }
        if (out->rowSatds) { CU_TRY(c, cudaMemcpyAsync(out->rowSatds, o.rowSatds, (size_t)g.hCU * 4, cudaMemcpyDeviceToHost, st)); c->stats.d2hBytes += g.hCU * 4; }
        if (!dBatchSums) CU_TRY(c, cudaMemcpyAsync(sums, c->dSmall, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    }
    return X265CU_OK;
}

/* lowresIntraEstimate of a few frames in one launch (blockIdx.y = frame); host destinations in mapped pinned memory are
 * written by one small scatter launch, others by copies.  dBatchSums: [count][2] u64, zeroed by the caller. */
static int intraEnqueueBatch(x265cu_ctx* c, int count, const int* slots, x265cu_intra_out* const* outs, unsigned long long* dBatchSums, cudaStream_t st)
{
    if (!st) st = c->stream;
    const GeomDev& g = c->g;
    for (int first = 0; first < count; first += INTRA_BATCH)
    {
        const int m = count - first < INTRA_BATCH ? count - first : INTRA_BATCH;
        IntraBatch ib;
        memset(&ib, 0, sizeof(ib));
        for (int k = 0; k < m; k++)
        {
            const int slot = slots[first + k];
            if (badSlot(c, slot)) return fail(c, X265CU_EINVAL, "x265cu_intra: bad slot");
            IntraOutDev& o = ib.o[k];
            ib.plane0[k] = slotPlane0(c, slot);
            o.intraCost = slotIntraCost(c, slot);
            o.intraMode = slotIntraMode(c, slot);
            o.lowresCosts = slotLowresCosts(c, slot, 0, 0);
            o.rowSatds = slotRowSatds(c, slot, 0, 0);
            o.sums = dBatchSums + (size_t)(first + k) * 2;
            o.invQ = c->hasInvQ[slot] ? slotInvQ(c, slot) : NULL;
            CU_TRY(c, cudaMemsetAsync(o.rowSatds, 0, (size_t)g.hCU * sizeof(int), st));
        }
        {
            KernelScope ks(c, X265CU_K_INTRA, 1, st);
            dim3 grid((g.nCU + 7) / 8, m);
            if (c->pb == 1) intra_batch_kernel<uint8_t><<<grid, 256, 0, st>>>(ib, g, c->cfg.lookaheadLambda, c->pixelMax);
            else intra_batch_kernel<uint16_t><<<grid, 256, 0, st>>>(ib, g, c->cfg.lookaheadLambda, c->pixelMax);
        }
        CU_TRY(c, cudaGetLastError());
        ScatterBatch sb;
        int ns = 0;
        for (int k = 0; k < m; k++)
        {
            x265cu_intra_out* out = outs[first + k];
            if (!out) continue;
            const IntraOutDev& o = ib.o[k];
            void* dst[4] = { out->intraCost, out->intraMode, out->lowresCosts, out->rowSatds };
            const void* src[4] = { o.intraCost, o.intraMode, o.lowresCosts, o.rowSatds };
            const size_t len[4] = { (size_t)g.nCU * 4, (size_t)g.nCU, (size_t)g.nCU * 2, (size_t)g.hCU * 4 };
            for (int a = 0; a < 4; a++)
            {
                if (!dst[a]) continue;
                uint8_t* alias = c->mappedResults ? mappedAlias(dst[a], len[a]) : NULL;
                if (alias)
                {
                    ScatterDev e = { (const uint8_t*)src[a], alias, (unsigned)len[a], 0 };
                    sb.e[ns++] = e;
                    if (ns == SCATTER_SMALL)
                    {
                        KernelScope ks(c, X265CU_K_RESULTS, 1, st);
                        scatter_small_kernel<<<ns, 256, 0, st>>>(sb);
                        CU_TRY(c, cudaGetLastError());
                        ns = 0;
                    }
                }
                else
                    CU_TRY(c, cudaMemcpyAsync(dst[a], src[a], len[a], cudaMemcpyDeviceToHost, st));
                c->stats.d2hBytes += (int64_t)len[a];
            }
        }
        if (ns)
        {
            KernelScope ks(c, X265CU_K_RESULTS, 1, st);
            scatter_small_kernel<<<ns, 256, 0, st>>>(sb);
            CU_TRY(c, cudaGetLastError());
        }
    }
    return X265CU_OK;
}

int x265cu_intra(x265cu_ctx* c, int slot, x265cu_intra_out* out)
{
    if (!c) return X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    unsigned long long sums[2] = { 0, 0 };
    int r = intraEnqueue(c, slot, out, sums, NULL, NULL);
    if (r) return r;
    r = syncStream(c);
    if (out) { out->sums[0] = (int64_t)sums[0]; out->sums[1] = (int64_t)sums[1]; }
    return r;
}

int x265cu_intra_batch(x265cu_ctx* c, int n, const int* slots, x265cu_intra_out* outs)
{
    if (!c || n < 0 || (n && (!slots || !outs))) return c ? fail(c, X265CU_EINVAL, "x265cu_intra_batch: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    if (growHost(c, &c->hPre, &c->hPreCap, (size_t)n * 16) || growDevice(c, &c->dPre, &c->dPreCap, (size_t)n * 16)) return X265CU_ECUDA;
    CU_TRY(c, cudaMemsetAsync(c->dPre, 0, (size_t)n * 16, c->stream));
    {
        std::vector<x265cu_intra_out*> po((size_t)n);
        for (int i = 0; i < n; i++) po[i] = &outs[i];
        int r = intraEnqueueBatch(c, n, slots, &po[0], (unsigned long long*)c->dPre, NULL);
        if (r) { cudaStreamSynchronize(c->stream); return r; }
    }
    CU_TRY(c, cudaMemcpyAsync(c->hPre, c->dPre, (size_t)n * 16, cudaMemcpyDeviceToHost, c->stream));
    int r = syncStream(c);
    if (r) return r;
    for (int i = 0; i < n; i++)
    {
        const unsigned long long* sm = (const unsigned long long*)(c->hPre + (size_t)i * 16);
        outs[i].sums[0] = (int64_t)sm[0];
        outs[i].sums[1] = (int64_t)sm[1];
    }
    return X265CU_OK;
}

/* ---- explicit weighted-prediction analysis (x265cu_wp.cuh; encoder/weightPrediction.cpp) ---- */
int x265cu_wp_prepare(x265cu_ctx* c, int fencSlot, int refSlot, int plane, const void* lowresMvs, const int32_t* intraCost)
{
    if (!c || badSlot(c, fencSlot) || badSlot(c, refSlot) || plane < 0 || plane > 2 || (plane == 0 && !intraCost))
        return c ? fail(c, X265CU_EINVAL, "x265cu_wp_prepare: bad argument") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    c->wpReady = false;
    const GeomDev& g = c->g;
    if (plane && (!c->hasChroma[fencSlot] || !c->hasChroma[refSlot]))
        return fail(c, X265CU_EINVAL, "x265cu_wp_prepare: the chroma planes of these frames are not on the device (frames must go through x265cu_frame_init_var* with chroma)");
    const size_t nCU = (size_t)g.nCU;
    const size_t lumaMc = (size_t)g.stride * g.lines * c->pb, chromaMc = (size_t)c->chromaPitch * c->chromaRows * c->pb;
    const size_t offIntra = alignUp(nCU * 4, 256), offMc = offIntra + alignUp(nCU * 4, 256);
    const size_t offCand = offMc + alignUp(lumaMc > chromaMc ? lumaMc : chromaMc, 256);
    if (growDevice(c, &c->dWp, &c->dWpCap, offCand + 64 * 1024)) return X265CU_ECUDA;
    int* dMvs = (int*)c->dWp;
    int* dIntra = (int*)(c->dWp + offIntra);
    uint8_t* dMc = c->dWp + offMc;
    if (lowresMvs)
    {
        CU_TRY(c, cudaMemcpyAsync(dMvs, lowresMvs, nCU * 4, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2dBytes += (int64_t)nCU * 4;
    }
    WpCostArgs& a = c->wpArgs;
    if (plane == 0)
    {
        /* the host's intraCost is the truth (rate control may have rescaled it since the lookahead wrote it) */
        CU_TRY(c, cudaMemcpyAsync(dIntra, intraCost, nCU * 4, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2dBytes += (int64_t)nCU * 4;
        a.fenc = slotPlane0(c, fencSlot); a.fencStride = g.stride;
        a.ref = slotPlane0(c, refSlot); a.refStride = g.stride;
        a.wBlk = g.width >> 3; a.nBlk = (g.width >> 3) * (g.lines >> 3);
        a.intraCost = dIntra;
        if (lowresMvs)
        {
            KernelScope ks(c, X265CU_K_WEIGHT);
            const int blocks = (g.nCU + 63) / 64;
            if (c->pb == 1) wp_mc_luma_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>((const uint8_t*)slotPlane0(c, refSlot), g, dMvs, (uint8_t*)dMc);
            else wp_mc_luma_kernel<uint16_t><<<blocks, 256, 0, c->stream>>>((const uint16_t*)slotPlane0(c, refSlot), g, dMvs, (uint16_t*)dMc);
            CU_TRY(c, cudaGetLastError());
            a.ref = dMc;
        }
    }
    else
    {
        /* the analysis area of the chroma planes: whole 8x8 blocks only (weightPrediction.cpp:329-336) */
        const int width = ((c->cfg.srcWidth >> 4) << 4) >> 1, height = ((c->cfg.srcHeight >> 4) << 4) >> 1;
        a.fenc = slotChroma(c, fencSlot, plane - 1); a.fencStride = c->chromaPitch;
        a.ref = slotChroma(c, refSlot, plane - 1); a.refStride = c->chromaPitch;
        a.wBlk = width >> 3; a.nBlk = (width >> 3) * (height >> 3);
        a.intraCost = NULL;
        if (lowresMvs && width > 0 && height > 0)
        {
            KernelScope ks(c, X265CU_K_WEIGHT);
            dim3 grid((width + 31) / 32, (height + 7) / 8);
            const int cW = c->cfg.srcWidth >> 1, cH = c->cfg.srcHeight >> 1;
            /* cache.lowresWidthInCU / HeightInCU = Lowres::width >> 3, lines >> 3 (weightPrediction.cpp:233-234) */
            if (c->pb == 1)
                wp_mc_chroma_kernel<uint8_t><<<grid, 256, 0, c->stream>>>((const uint8_t*)a.ref, c->chromaPitch, cW, cH, dMvs, g.width >> 3, g.lines >> 3, width, height,
                                                                         (uint8_t*)dMc, c->chromaPitch, c->cfg.bitDepth);
            else
                wp_mc_chroma_kernel<uint16_t><<<grid, 256, 0, c->stream>>>((const uint16_t*)a.ref, c->chromaPitch, cW, cH, dMvs, g.width >> 3, g.lines >> 3, width, height,
                                                                          (uint16_t*)dMc, c->chromaPitch, c->cfg.bitDepth);
            CU_TRY(c, cudaGetLastError());
            a.ref = dMc;
        }
    }
    c->wpReady = true;
    return X265CU_OK;
}

int x265cu_wp_cost(x265cu_ctx* c, int n, const x265cu_weight_item* cands, uint32_t* costs)
{
    if (!c || n < 0 || n > 1024 || (n && (!cands || !costs))) return c ? fail(c, X265CU_EINVAL, "x265cu_wp_cost: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    if (!c->wpReady) return fail(c, X265CU_EINVAL, "x265cu_wp_cost: no x265cu_wp_prepare before it");
    const GeomDev& g = c->g;
    const size_t nCU = (size_t)g.nCU;
    const size_t lumaMc = (size_t)g.stride * g.lines * c->pb, chromaMc = (size_t)c->chromaPitch * c->chromaRows * c->pb;
    const size_t offCand = alignUp(nCU * 4, 256) * 2 + alignUp(lumaMc > chromaMc ? lumaMc : chromaMc, 256);
    WpCand h[1024];
    for (int i = 0; i < n; i++)
    {
        h[i].weighted = cands[i].weighted; h[i].scale = cands[i].scale;
        weightArgs(c, cands[i].scale, cands[i].denom, cands[i].offset, &h[i].round, &h[i].shift, &h[i].offset);
    }
    WpCand* dCand = (WpCand*)(c->dWp + offCand);
    unsigned int* dCosts = (unsigned int*)(c->dWp + offCand + alignUp((size_t)n * sizeof(WpCand), 256));
    CU_TRY(c, cudaMemcpyAsync(dCand, h, (size_t)n * sizeof(WpCand), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemsetAsync(dCosts, 0, (size_t)n * sizeof(unsigned int), c->stream));
    {
        KernelScope ks(c, X265CU_K_WEIGHT);
        int bx = (c->wpArgs.nBlk / 8 + 7) / 8;
        if (bx > 148 * 2) bx = 148 * 2;
        if (bx < 1) bx = 1;
        dim3 grid(bx, n);
        if (c->pb == 1) wp_cost_kernel<uint8_t><<<grid, 256, 0, c->stream>>>(c->wpArgs, dCand, dCosts, c->correction, c->pixelMax);
        else wp_cost_kernel<uint16_t><<<grid, 256, 0, c->stream>>>(c->wpArgs, dCand, dCosts, c->correction, c->pixelMax);
        CU_TRY(c, cudaGetLastError());
    }
    CU_TRY(c, cudaMemcpyAsync(costs, dCosts, (size_t)n * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2hBytes += (int64_t)n * 4;
    return syncStream(c);
}

/* ---- cuTree propagation (x265cu_cutree.cuh) ---- */
int x265cu_frame_set_propagate(x265cu_ctx* c, int slot, const uint16_t* propagateCost)
{
    if (!c || badSlot(c, slot) || !propagateCost) return c ? fail(c, X265CU_EINVAL, "x265cu_frame_set_propagate: bad argument") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const size_t n = (size_t)c->g.nCU;
    if (growHost(c, (uint8_t**)&c->hPropOut, &c->hPropOutCap, n * sizeof(unsigned long long))) return X265CU_ECUDA;
    /* the staging area is reused: earlier copies out of / into it must have finished */
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    unsigned long long* w = (unsigned long long*)c->hPropOut;
    for (size_t i = 0; i < n; i++) w[i] = propagateCost[i];
    CU_TRY(c, cudaMemcpyAsync(c->dPropagate + (size_t)slot * n, w, n * sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stats.h2dBytes += (int64_t)(n * sizeof(unsigned long long));
    return X265CU_OK;
}

int x265cu_frame_set_array(x265cu_ctx* c, int slot, int which, int d0, int d1, const void* data)
{
    if (!c || badSlot(c, slot) || !data) return c ? fail(c, X265CU_EINVAL, "x265cu_frame_set_array: bad argument") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const size_t n = (size_t)c->g.nCU;
    void* dst; size_t bytes;
    if (which == 4 && d0 >= 0 && d0 < c->bf + 2 && d1 >= 0 && d1 < c->bf + 2) { dst = slotLowresCosts(c, slot, d0, d1); bytes = n * sizeof(uint16_t); }
    else if (which == 6 && (d0 == 0 || d0 == 1) && d1 >= 1 && d1 <= c->bf + 1) { dst = slotMvs(c, slot, d0, d1); bytes = n * sizeof(int); }
    else return fail(c, X265CU_EINVAL, "x265cu_frame_set_array: unknown array");
    /* stream-ordered before the kernels that read it; waits, so the caller's array may change afterwards */
    CU_TRY(c, cudaMemcpyAsync(dst, data, bytes, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaStreamSynchronize(c->stream));
    c->stats.h2dBytes += (int64_t)bytes;
    return X265CU_OK;
}

int x265cu_cutree_run(x265cu_ctx* c, int n, const x265cu_cutree_op* ops, int nOut, const int* outSlots, uint16_t* const* outPropagateCost)
{
    if (c) c->lastEnqueued = "cutree_run";
    SlowLog slow("x265cu_cutree_run");
    if (!c || n < 0 || nOut < 0 || (n && !ops) || (nOut && (!outSlots || !outPropagateCost)))
        return c ? fail(c, X265CU_EINVAL, "x265cu_cutree_run: bad argument") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const int bf = c->bf;
    const size_t nCU = (size_t)c->g.nCU;
    for (int i = 0; i < n; i++)
    {
        const x265cu_cutree_op& o = ops[i];
        if (badSlot(c, o.fenc)) return fail(c, X265CU_EINVAL, "x265cu_cutree_run: bad frame slot");
        if (o.kind == X265CU_CT_ZERO) continue;
        if (o.kind != X265CU_CT_PROPAGATE) return fail(c, X265CU_EINVAL, "x265cu_cutree_run: unknown op");
        if (badSlot(c, o.ref0) || badSlot(c, o.ref1) || o.d0 < 1 || o.d0 > bf + 1 || o.d1 < 0 || o.d1 > bf + 1)
            return fail(c, X265CU_EINVAL, "x265cu_cutree_run: bad references");
        if (!c->hasInvQ[o.fenc]) return fail(c, X265CU_EINVAL, "x265cu_cutree_run: frame has no invQscaleFactor (cuTree needs the AQ arrays)");
    }
    for (int i = 0; i < nOut; i++)
        if (badSlot(c, outSlots[i]) || !outPropagateCost[i]) return fail(c, X265CU_EINVAL, "x265cu_cutree_run: bad output");
    if (n + nOut == 0) return X265CU_OK;
    if (nOut)
    {
        if (growDevice(c, (uint8_t**)&c->dPropOut, &c->dPropOutCap, (size_t)nOut * nCU * sizeof(uint16_t))) return X265CU_ECUDA;
        if (growHost(c, (uint8_t**)&c->hPropOut, &c->hPropOutCap, (size_t)nOut * nCU * sizeof(unsigned long long))) return X265CU_ECUDA;
    }

    CutreeArgs a;
    a.wCU = c->g.wCU; a.hCU = c->g.hCU; a.nCU = c->g.nCU;
    a.costTables = (bf + 2) * (bf + 2); a.mvFields = 2 * (bf + 1);
    a.intraCost = c->dIntraCost; a.invQ = c->dInvQ; a.lowresCosts = c->dLowresCosts; a.mvs = c->dMvs;
    a.acc = c->dPropagate; a.out = c->dPropOut;
    a.nOps = 0;

    /* cooperative launch: one CTA per SM, all co-resident (grid barrier between the phases) */
    if (c->cutreeCtas <= 0)
    {
        int sms = 0, perSm = 0;
        CU_TRY(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device));
        CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, cutree_kernel, CUTREE_THREADS, 0));
        if (sms < 1 || perSm < 1) return fail(c, X265CU_ECUDA, "x265cu_cutree_run: the cuTree kernel does not fit an SM");
        c->cutreeCtas = sms;
        if (const char* e = getenv("X265CU_CUTREE_CTAS")) { int v = atoi(e); if (v >= 1 && v <= sms * perSm) c->cutreeCtas = v; }
    }
    cudaLaunchConfig_t lc;
    memset(&lc, 0, sizeof(lc));
    lc.gridDim = dim3(c->cutreeCtas); lc.blockDim = dim3(CUTREE_THREADS); lc.dynamicSmemBytes = 0; lc.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = 1;
    lc.attrs = attr; lc.numAttrs = 1;

    {
        KernelScope ks(c, X265CU_K_CUTREE, 0);
        /* ops in the caller's order, then one PACK per requested output; CUTREE_MAX_OPS per launch (launches of
         * one stream run in order, so a longer list simply continues in the next launch) */
        for (int i = 0; i < n + nOut; i++)
        {
            CutreeOpDev& d = a.ops[a.nOps++];
            memset(&d, 0, sizeof(d));
            if (i < n)
            {
                const x265cu_cutree_op& o = ops[i];
                d.kind = o.kind == X265CU_CT_ZERO ? CT_OP_ZERO : CT_OP_PROPAGATE;
                d.fenc = o.fenc;
                if (d.kind == CT_OP_PROPAGATE)
                {
                    d.ref0 = o.ref0; d.ref1 = o.ref1;
                    d.costOfs = o.d0 * (bf + 2) + o.d1;
                    d.mvOfs0 = o.d0 - 1;
                    d.mvOfs1 = o.d1 > 0 ? (bf + 1) + o.d1 - 1 : -1;
                    d.referenced = o.referenced != 0;
                    d.bipredWeight = o.bipredWeight;
                    d.fps = o.fpsFactor * (1.0 / 256);      /* exact scaling, as the reference's `*fpsFactor / 256` */
                }
            }
            else
            {
                d.kind = CT_OP_PACK; d.fenc = outSlots[i - n]; d.outIndex = i - n;
            }
            if (a.nOps == CUTREE_MAX_OPS || i == n + nOut - 1)
            {
                cutree_schedule(a.ops, a.nOps);
                CU_TRY(c, cudaLaunchKernelEx(&lc, cutree_kernel, a));
                c->stats.launches[X265CU_K_CUTREE]++;
                a.nOps = 0;
            }
        }
    }
    if (nOut)
    {
        CU_TRY(c, cudaMemcpyAsync(c->hPropOut, c->dPropOut, (size_t)nOut * nCU * sizeof(uint16_t), cudaMemcpyDeviceToHost, c->stream));
        if (syncStream(c)) return X265CU_ECUDA;
        for (int i = 0; i < nOut; i++)
            memcpy(outPropagateCost[i], c->hPropOut + (size_t)i * nCU, nCU * sizeof(uint16_t));
        c->stats.d2hBytes += (int64_t)((size_t)nOut * nCU * sizeof(uint16_t));
    }
    return X265CU_OK;
}

/* -------------------------------------------------------------------------------------------- */
static void weightArgs(const x265cu_ctx* c, int scale, int denom, int offset, int* round, int* shift, int* off)
{
    /* weightCostLuma / weightsAnalyse call sites, slicetype.cpp:345-352,476-484 */
    int r = denom ? 1 << (denom - 1) : 0;
    *round = r << c->correction;
    *shift = denom + c->correction;
    *off = offset << (c->cfg.bitDepth - 8);
    (void)scale;
}

int x265cu_weight_cost_batch(x265cu_ctx* c, int n, const x265cu_weight_item* items, uint32_t* costs)
{
    if (!c || n < 0 || (n && (!items || !costs))) return c ? fail(c, X265CU_EINVAL, "x265cu_weight_cost_batch: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const GeomDev& g = c->g;
    size_t argBytes = alignUp((size_t)n * sizeof(WeightCostDev), 256) + (size_t)n * sizeof(unsigned int);
    if (growHost(c, &c->hArgs, &c->hArgsCap, argBytes) || growDevice(c, &c->dArgs, &c->dArgsCap, argBytes)) return X265CU_ECUDA;
    WeightCostDev* h = (WeightCostDev*)c->hArgs;
    for (int i = 0; i < n; i++)
    {
        if (badSlot(c, items[i].fenc) || badSlot(c, items[i].ref)) return fail(c, X265CU_EINVAL, "x265cu_weight_cost_batch: bad slot");
        h[i].fenc = slotPlane0(c, items[i].fenc);
        h[i].ref = slotPlane0(c, items[i].ref);
        h[i].intraCost = slotIntraCost(c, items[i].fenc);
        h[i].weighted = items[i].weighted;
        h[i].scale = items[i].scale;
        weightArgs(c, items[i].scale, items[i].denom, items[i].offset, &h[i].round, &h[i].shift, &h[i].offset);
    }
    unsigned int* dCosts = (unsigned int*)(c->dArgs + alignUp((size_t)n * sizeof(WeightCostDev), 256));
    CU_TRY(c, cudaMemcpyAsync(c->dArgs, h, (size_t)n * sizeof(WeightCostDev), cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemsetAsync(dCosts, 0, (size_t)n * sizeof(unsigned int), c->stream));
    {
        KernelScope ks(c, X265CU_K_WEIGHT);
        int bx = (g.nCU / 8 + 7) / 8;
        if (bx > 148 * 2) bx = 148 * 2;
        if (bx < 1) bx = 1;
        dim3 grid(bx, n);
        if (c->pb == 1)
            weight_cost_kernel<uint8_t><<<grid, 256, 0, c->stream>>>((const WeightCostDev*)c->dArgs, dCosts, g, c->correction, c->pixelMax);
        else
            weight_cost_kernel<uint16_t><<<grid, 256, 0, c->stream>>>((const WeightCostDev*)c->dArgs, dCosts, g, c->correction, c->pixelMax);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(costs, dCosts, (size_t)n * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2hBytes += (int64_t)n * 4;
    return syncStream(c);
}

/* -------------------------------------------------------------------------------------------- */
int x265cu_estimate_batch(x265cu_ctx* c, int n, const x265cu_job* jobs, x265cu_job_result* results)
{
    if (c) c->lastEnqueued = "estimate_batch";
    SlowLog slow("x265cu_estimate_batch");
    if (!c || n < 0 || (n && (!jobs || !results))) return c ? fail(c, X265CU_EINVAL, "x265cu_estimate_batch: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const GeomDev& g = c->g;
    const size_t nCU = (size_t)g.nCU, hCU = (size_t)g.hCU;
    const std::chrono::steady_clock::time_point tH0 = std::chrono::steady_clock::now();

    /* ---- plan: per-job packed record + one SearchPlan per (job, list) searched ---- */
    std::vector<size_t> recOff(n);
    std::vector<SearchItem> items;
    std::vector<SearchPlan> plans;
    std::vector<int> costIdx;
    std::vector<int> weightedJobs;
    /* staging layout: [n x 32-byte sums][record 0][record 1]...; when the caller wants no arrays
     * back (all destination pointers NULL) only the sums cross PCIe */
    const size_t sumsBytes = alignUp((size_t)n * 32, 256);
    size_t total = sumsBytes;
    bool wantArrays = false;
    for (int i = 0; i < n; i++)
    {
        const x265cu_job& j = jobs[i];
        if (badSlot(c, j.fenc) || badSlot(c, j.ref0) || (j.d1 > 0 && badSlot(c, j.ref1)) || j.d0 < 1 || j.d0 > c->bf + 1 || j.d1 < 0 || j.d1 > c->bf + 1)
            return fail(c, X265CU_EINVAL, "x265cu_estimate_batch: bad job");
        if (j.doSearch[1] && j.d1 == 0) return fail(c, X265CU_EINVAL, "x265cu_estimate_batch: L1 search without p1");
        recOff[i] = total;
        if (j.mvs[0] || j.mvs[1] || j.mvCosts[0] || j.mvCosts[1] || j.lowresCosts || j.rowSatds) wantArrays = true;
        size_t rec = alignUp(hCU * 4, 16) + alignUp(nCU * 2, 16);
        if (j.doSearch[0]) rec += 2 * alignUp(nCU * 4, 16);
        if (j.doSearch[1]) rec += 2 * alignUp(nCU * 4, 16);
        total += rec;
        const bool useSlices = j.sliced && c->cfg.numCoopSlices > 1;   /* (p1 > b || search) holds for a searching job */
        for (int l = 0; l < 2; l++)
        {
            if (!j.doSearch[l]) continue;
            /* the two lists of an estimate are independent searches */
            SearchPlan pl;
            memset(&pl, 0, sizeof(pl));
            pl.job = i;
            pl.list = l;
            pl.rowsPerSlice = useSlices ? c->cfg.numRowsPerSlice : g.hCU;
            pl.numSlices = useSlices ? c->cfg.numCoopSlices : 1;
            plans.push_back(pl);
        }
        costIdx.push_back(i);     /* bidir / intra decision + sums of every estimate: cost_kernel */
        if (j.weighted && j.doSearch[0]) weightedJobs.push_back(i);
    }
    /* weighted reference pool */
    while (c->wPool.size() < weightedJobs.size())
    {
        const size_t wBytes = (size_t)4 * g.planeSize * c->pb + 256;
        size_t got = 0;
        void* p = poolTake(g_devPool, c->cfg.device, wBytes, &got);
        if (p && got != wBytes) { poolGive(g_devPool, c->cfg.device, p, got); p = NULL; }    /* only an area of exactly this use */
        if (!p) CU_TRY(c, cudaMalloc(&p, wBytes));
        c->wPool.push_back(p);
    }

    const std::chrono::steady_clock::time_point tP1 = std::chrono::steady_clock::now();
    /* ---- hints for the speculation kernel (speed only, never a result): a finished MV field of the
     * temporally nearest frame for the same list and distance.  Searches of a batch that have no such
     * field yet are run in two WAVES: one seed per (list, distance) first, the others after it with the
     * seed's field as their hint. ---- */
    std::vector<int> classOf(plans.size(), 0);   /* 0/1: speculative path, wave 0/1; 2: plain kernel */
    if (!plans.empty())
    {
        for (;;)
        {
            /* relative picture order of the slots, propagated through the jobs' (fenc, ref0, ref1, d0, d1) relations */
            for (int pass = 0; pass < n + 1; pass++)
            {
                bool changed = false;
                for (int i = 0; i < n; i++)
                {
                    const x265cu_job& j = jobs[i];
                    const int F = j.fenc, R0 = j.ref0, R1 = j.d1 > 0 ? j.ref1 : -1;
                    if (!c->slotPocKnown[F])
                    {
                        if (c->slotPocKnown[R0]) { c->slotPoc[F] = c->slotPoc[R0] + j.d0; c->slotPocKnown[F] = 1; changed = true; }
                        else if (R1 >= 0 && c->slotPocKnown[R1]) { c->slotPoc[F] = c->slotPoc[R1] - j.d1; c->slotPocKnown[F] = 1; changed = true; }
                    }
                    if (c->slotPocKnown[F])
                    {
                        if (!c->slotPocKnown[R0]) { c->slotPoc[R0] = c->slotPoc[F] - j.d0; c->slotPocKnown[R0] = 1; changed = true; }
                        if (R1 >= 0 && !c->slotPocKnown[R1]) { c->slotPoc[R1] = c->slotPoc[F] + j.d1; c->slotPocKnown[R1] = 1; changed = true; }
                    }
                }
                if (!changed) break;
            }
            /* a job none of whose frames is placed yet (a new stream of slots): give it an origin of its own */
            int orphan = -1;
            for (int i = 0; i < n && orphan < 0; i++)
                if (!c->slotPocKnown[jobs[i].fenc]) orphan = i;
            if (orphan < 0) break;
            c->pocBase += 1 << 20;
            c->slotPoc[jobs[orphan].fenc] = c->pocBase;
            c->slotPocKnown[jobs[orphan].fenc] = 1;
        }
        const size_t perSlot = (size_t)2 * (c->bf + 1);
        const int maxStep = 8;
        std::vector<int> hintDist(plans.size(), 0);
        for (size_t k = 0; k < plans.size(); k++)
        {
            SearchPlan& pl = plans[k];
            const x265cu_job& j = jobs[pl.job];
            const int F = j.fenc, d = pl.list ? j.d1 : j.d0;
            long long best = maxStep + 1;
            for (size_t sl = 0; sl < c->slotPocKnown.size(); sl++)
            {
                if ((int)sl == F || !c->slotPocKnown[sl] || !c->mvValid[sl * perSlot + (size_t)pl.list * (c->bf + 1) + (d - 1)]) continue;
                long long dist = c->slotPoc[sl] - c->slotPoc[F];
                if (dist < 0) dist = -dist;
                if (dist < best) { best = dist; pl.hint = slotMvs(c, (int)sl, pl.list, d); hintDist[k] = (int)dist; }
            }
        }
        /* Which kernel runs a search.  The speculative path (refine + commit) does more work per CU to shorten the
         * dependent chain, which pays when the GPU is mostly idle (few searches in flight) AND the hint is close in
         * time; a batch that fills the GPU, or a search without a good hint, takes the plain wavefront kernel. */
        for (size_t k = 0; k < plans.size(); k++)
        {
            if (c->searchMode == 0) classOf[k] = 2;
            else if (c->searchMode == 1) classOf[k] = 0;
            else classOf[k] = (plans[k].hint && hintDist[k] <= c->specMaxDist && (int)plans.size() <= c->specMaxPlans) ? 0 : 2;
            if (classOf[k] == 2) plans[k].hint = NULL;
        }
        /* (experiments, searchMode 1) in-batch seeds: per (list, distance) class still without a hint, the plan in the
         * middle of the class's picture-order range goes first (wave 0, unhinted), the rest follow in wave 1 */
        for (int l = 0; l < 2 && c->searchMode == 1; l++)
            for (int d = 1; d <= c->bf + 1; d++)
            {
                std::vector<size_t> cls;
                for (size_t k = 0; k < plans.size(); k++)
                {
                    const x265cu_job& j = jobs[plans[k].job];
                    if (!plans[k].hint && plans[k].list == l && (l ? j.d1 : j.d0) == d) cls.push_back(k);
                }
                if (cls.size() < 2) continue;
                std::sort(cls.begin(), cls.end(), [&](size_t a, size_t b) { return c->slotPoc[jobs[plans[a].job].fenc] < c->slotPoc[jobs[plans[b].job].fenc]; });
                const size_t seed = cls[cls.size() / 2];
                const int* seedField = slotMvs(c, jobs[plans[seed].job].fenc, l, d);
                for (size_t i = 0; i < cls.size(); i++)
                    if (cls[i] != seed) { plans[cls[i]].hint = seedField; classOf[cls[i]] = 1; }
            }
#ifdef X265CU_SEARCH_STATS
        for (size_t k = 0; k < plans.size(); k++) { c->dbgPlans[0]++; if (classOf[k] != 2) c->dbgPlans[classOf[k] ? 2 : 1]++; }
#endif
        /* fields this batch produces become hints for later batches */
        for (size_t k = 0; k < plans.size(); k++)
        {
            const x265cu_job& j = jobs[plans[k].job];
            c->mvValid[(size_t)j.fenc * perSlot + (size_t)plans[k].list * (c->bf + 1) + ((plans[k].list ? j.d1 : j.d0) - 1)] = 1;
        }
        /* order: speculative wave 0, speculative wave 1, plain */
        {
            std::vector<SearchPlan> sorted;
            for (int w = 0; w < 3; w++)
                for (size_t k = 0; k < plans.size(); k++)
                    if (classOf[k] == w) sorted.push_back(plans[k]);
            plans.swap(sorted);
        }
    }
    size_t classPlans[4] = { 0, 0, 0, plans.size() };       /* first plan of each class */
    for (size_t k = 0; k < classOf.size(); k++) { if (classOf[k] < 1) classPlans[1]++; if (classOf[k] < 2) classPlans[2]++; }
    const size_t memoPerSearch = nCU * MEMO_N * sizeof(int4) + 3 * alignUp(nCU * sizeof(int), 256);   /* memo + the three estimate fields */
    if (classPlans[2] && growDevice(c, &c->dMemo, &c->dMemoCap, classPlans[2] * memoPerSearch)) return X265CU_ECUDA;

    const std::chrono::steady_clock::time_point tP2 = std::chrono::steady_clock::now();
    /* ---- commit work items: one per row group of every (search, cooperative slice) ---- */
    int maxItemRows = 1, maxPlainRows = 1;
    int handRows = 0;            /* global hand-off rows (one per row group that has a group above it) */
    size_t classItems[4] = { 0, 0, 0, 0 };
    for (size_t k = 0; k < plans.size(); k++)
    {
        SearchPlan& pl = plans[k];
        const bool plain = k >= classPlans[2];
        if (!plain)
        {
            pl.memo = (int4*)(c->dMemo + k * memoPerSearch);
            pl.field[0] = (int*)(c->dMemo + k * memoPerSearch + nCU * MEMO_N * sizeof(int4));
            pl.field[1] = pl.field[0] + alignUp(nCU * sizeof(int), 256) / sizeof(int);
            pl.field[2] = pl.field[1] + alignUp(nCU * sizeof(int), 256) / sizeof(int);
        }
        if (k == classPlans[1]) classItems[1] = items.size();
        if (k == classPlans[2]) classItems[2] = items.size();
        const int groupRows = plain ? (c->plainOct ? 4 * c->octWarps : c->plainWarps) : c->searchWarps;
        for (int s = 0; s < pl.numSlices; s++)
        {
            const int sFirst = pl.numSlices > 1 ? pl.rowsPerSlice * s : 0;
            const int sLast = s == pl.numSlices - 1 ? g.hCU - 1 : pl.rowsPerSlice * (s + 1) - 1;
            /* row groups, bottom first: a group only waits for a group with a lower block index */
            int prevPub = -1;
            for (int bottom = sLast; bottom >= sFirst; bottom -= groupRows)
            {
                SearchItem it;
                it.search = (int)k;
                it.sliceFirstY = sFirst;
                it.sliceLastY = sLast;
                it.lastY = bottom;
                it.firstY = bottom - groupRows + 1 > sFirst ? bottom - groupRows + 1 : sFirst;
                it.subBase = prevPub;                                   /* hand-off row of the group below */
                it.pubBase = it.firstY > sFirst ? handRows++ * g.wCU : -1; /* the top group has nobody above */
                prevPub = it.pubBase;
                int& mr = plain ? maxPlainRows : maxItemRows;
                if (it.lastY - it.firstY + 1 > mr) mr = it.lastY - it.firstY + 1;
                items.push_back(it);
            }
        }
    }
    if (classPlans[1] == plans.size()) classItems[1] = items.size();
    if (classPlans[2] == plans.size()) classItems[2] = items.size();
    /* Launch order of the plain kernel's CTAs.  A group becomes useful once the group below it is a few CUs ahead, so
     * with the groups of one search in consecutive blocks a batch that does not fit the GPU fills its CTA slots with the
     * upper groups of a few searches, most of them waiting for the wavefront to reach them.  Ordered by LEVEL (the
     * bottom groups of every search and slice first, then the second groups, ...) the resident CTAs are the ones whose
     * lower neighbour is running right now.  A group still only waits for a lower block index. */
    if (items.size() - classItems[2] > 1)
    {
        std::vector<std::pair<int, SearchItem> > keyed;
        for (size_t k = classItems[2]; k < items.size(); k++)
        {
            const SearchItem& it = items[k];
            keyed.push_back(std::make_pair((it.sliceLastY - it.lastY) / (c->plainOct ? 4 * c->octWarps : c->plainWarps), it));
        }
        std::stable_sort(keyed.begin(), keyed.end(), [](const std::pair<int, SearchItem>& a, const std::pair<int, SearchItem>& b) { return a.first < b.first; });
        for (size_t k = 0; k < keyed.size(); k++) items[classItems[2] + k] = keyed[k].second;
    }
    if (classPlans[1] == classPlans[2] && classPlans[2] < plans.size()) classItems[1] = classItems[2];
    classItems[3] = items.size();

    const std::chrono::steady_clock::time_point tP3 = std::chrono::steady_clock::now();
    size_t offJobs = 0;
    size_t offItems = alignUp(offJobs + (size_t)n * sizeof(JobDev), 256);
    size_t offPlans = alignUp(offItems + items.size() * sizeof(SearchItem), 256);
    size_t offCost = alignUp(offPlans + plans.size() * sizeof(SearchPlan), 256);
    size_t offW = alignUp(offCost + costIdx.size() * sizeof(int), 256);
    /* where each result array goes: straight into the caller's array when that lies in mapped pinned memory
     * (x265cu_host_register) -- a kernel writes it there over PCIe, nothing is staged or memcpy'd on the host --
     * otherwise through the pinned staging area and a host memcpy */
    struct HostCopy { void* dst; size_t srcOff; size_t bytes; };
    std::vector<HostCopy> hostCopies;
    std::vector<ScatterDev> devCopies;
    size_t devCopyBytes = 0;
    if (wantArrays)
    {
        for (int i = 0; i < n; i++)
        {
            const x265cu_job& j = jobs[i];
            size_t off = recOff[i];
            void* dsts[6]; size_t offs[6], lens[6]; int nd = 0;
            dsts[nd] = j.rowSatds; offs[nd] = off; lens[nd++] = hCU * 4; off += alignUp(hCU * 4, 16);
            dsts[nd] = j.lowresCosts; offs[nd] = off; lens[nd++] = nCU * 2; off += alignUp(nCU * 2, 16);
            for (int l = 0; l < 2; l++)
                if (j.doSearch[l])
                {
                    dsts[nd] = j.mvs[l]; offs[nd] = off; lens[nd++] = nCU * 4; off += alignUp(nCU * 4, 16);
                    dsts[nd] = j.mvCosts[l]; offs[nd] = off; lens[nd++] = nCU * 4; off += alignUp(nCU * 4, 16);
                }
            for (int k = 0; k < nd; k++)
            {
                if (!dsts[k]) continue;
                uint8_t* alias = c->mappedResults ? mappedAlias(dsts[k], lens[k]) : NULL;
                if (alias) { ScatterDev sd = { NULL, alias, (unsigned)lens[k], (unsigned)offs[k] }; devCopies.push_back(sd); devCopyBytes += lens[k]; }
                else { HostCopy hc = { dsts[k], offs[k], lens[k] }; hostCopies.push_back(hc); }
            }
        }
    }
    size_t offScat = alignUp(offW + weightedJobs.size() * sizeof(WeightDev), 256);
    size_t offProg = alignUp(offScat + devCopies.size() * sizeof(ScatterDev), 256);   /* device only: wavefront progress */
    size_t argBytes = alignUp(offProg + (size_t)handRows * g.wCU * sizeof(unsigned long long) + sizeof(SearchCtl), 256);
    if (growHost(c, &c->hArgs, &c->hArgsCap, argBytes) || growDevice(c, &c->dArgs, &c->dArgsCap, argBytes)) return X265CU_ECUDA;
    if (growDevice(c, &c->dStage, &c->dStageCap, total) || growHost(c, &c->hStage, &c->hStageCap, total)) return X265CU_ECUDA;

    JobDev* hj = (JobDev*)(c->hArgs + offJobs);
    std::vector<int> wSlotOfJob(n, -1);
    for (size_t k = 0; k < weightedJobs.size(); k++) wSlotOfJob[weightedJobs[k]] = (int)k;
    for (int i = 0; i < n; i++)
    {
        const x265cu_job& j = jobs[i];
        JobDev& d = hj[i];
        memset(&d, 0, sizeof(d));
        const int ref1 = j.d1 > 0 ? j.ref1 : j.fenc;
        d.fenc = slotPlane0(c, j.fenc);
        d.ref0 = slotPlane0(c, j.ref0);
        d.ref0w = wSlotOfJob[i] >= 0 ? (const void*)((uint8_t*)c->wPool[wSlotOfJob[i]] + (size_t)g.padOffset * c->pb) : d.ref0;
        d.ref1 = slotPlane0(c, ref1);
        d.mvs[0] = slotMvs(c, j.fenc, 0, j.d0);
        d.mvCosts[0] = slotMvCosts(c, j.fenc, 0, j.d0);
        d.mvs[1] = j.d1 > 0 ? slotMvs(c, j.fenc, 1, j.d1) : NULL;
        d.mvCosts[1] = j.d1 > 0 ? slotMvCosts(c, j.fenc, 1, j.d1) : NULL;
        d.lowresCosts = slotLowresCosts(c, j.fenc, j.d0, j.d1);
        d.rowSatds = slotRowSatds(c, j.fenc, j.d0, j.d1);
        d.intraCost = slotIntraCost(c, j.fenc);
        d.invQ = c->hasInvQ[j.fenc] ? slotInvQ(c, j.fenc) : NULL;
        uint8_t* rec = c->dStage + recOff[i];
        d.outSums = (unsigned long long*)(c->dStage + (size_t)i * 32);
        d.outRows = (int*)rec; rec += alignUp(hCU * 4, 16);
        d.outLowresCosts = (uint16_t*)rec; rec += alignUp(nCU * 2, 16);
        for (int l = 0; l < 2; l++)
            if (j.doSearch[l])
            {
                d.outMvs[l] = (int*)rec; rec += alignUp(nCU * 4, 16);
                d.outMvCosts[l] = (int*)rec; rec += alignUp(nCU * 4, 16);
            }
        d.d0 = j.d0; d.d1 = j.d1;
        d.doSearch[0] = j.doSearch[0] != 0; d.doSearch[1] = j.doSearch[1] != 0;
        d.bidir = j.d1 > 0;
        d.tmaZ[0] = wSlotOfJob[i] >= 0 ? -1 : 4 * j.ref0;     /* a weighted reference copy lives outside the mirrors */
        d.tmaZ[1] = 4 * ref1;
    }
    if (!items.empty()) memcpy(c->hArgs + offItems, &items[0], items.size() * sizeof(SearchItem));
    if (!plans.empty()) memcpy(c->hArgs + offPlans, &plans[0], plans.size() * sizeof(SearchPlan));
    if (!costIdx.empty()) memcpy(c->hArgs + offCost, &costIdx[0], costIdx.size() * sizeof(int));
    WeightDev* hw = (WeightDev*)(c->hArgs + offW);
    for (size_t k = 0; k < weightedJobs.size(); k++)
    {
        const x265cu_job& j = jobs[weightedJobs[k]];
        hw[k].src = slotBuffer(c, j.ref0);
        hw[k].dst = c->wPool[k];
        hw[k].scale = j.wScale;
        weightArgs(c, j.wScale, j.wDenom, j.wOffset, &hw[k].round, &hw[k].shift, &hw[k].offset);
    }
    {
        ScatterDev* hs = (ScatterDev*)(c->hArgs + offScat);
        for (size_t k = 0; k < devCopies.size(); k++) { hs[k] = devCopies[k]; hs[k].src = c->dStage + devCopies[k].srcOff; }
    }
    const std::chrono::steady_clock::time_point tH1 = std::chrono::steady_clock::now();
    CU_TRY(c, cudaMemcpyAsync(c->dArgs, c->hArgs, offProg, cudaMemcpyHostToDevice, c->stream));
    c->stats.h2dBytes += (int64_t)offProg;
    CU_TRY(c, cudaMemsetAsync(c->dStage, 0, sumsBytes, c->stream));
    if (!items.empty())
        CU_TRY(c, cudaMemsetAsync(c->dArgs + offProg, 0, (size_t)handRows * g.wCU * sizeof(unsigned long long) + sizeof(SearchCtl), c->stream));

    if (!weightedJobs.empty())
    {
        KernelScope ks(c, X265CU_K_WEIGHT);
        dim3 grid(148, (unsigned)weightedJobs.size());
        if (c->pb == 1)
            weight_planes_kernel<uint8_t><<<grid, 256, 0, c->stream>>>((const WeightDev*)(c->dArgs + offW), 4 * g.planeSize, c->correction, c->pixelMax);
        else
            weight_planes_kernel<uint16_t><<<grid, 256, 0, c->stream>>>((const WeightDev*)(c->dArgs + offW), 4 * g.planeSize, c->correction, c->pixelMax);
        CU_TRY(c, cudaGetLastError());
    }
    if (!items.empty())
    {
        /* kernels this batch launches for its searches: the plain kernel, and per speculative wave its refine iterations + commit */
        int nKernels = classItems[3] > classItems[2] ? 1 : 0;
        for (int w = 0; w < 2; w++)
        {
            const size_t np = classPlans[w + 1] - classPlans[w];
            if (np) nKernels += 1 + (np >= 32 ? (c->searchSpec < 1 ? c->searchSpec : 1) : c->searchSpec);
        }
        KernelScope ks(c, X265CU_K_SEARCH, nKernels);
        const JobDev* dJobs = (const JobDev*)(c->dArgs + offJobs);
        const SearchPlan* dPlans = (const SearchPlan*)(c->dArgs + offPlans);
        const SearchItem* dItems = (const SearchItem*)(c->dArgs + offItems);
        const uint16_t* dLutC = c->dLut + 2 * 32768;
        unsigned long long* dProg = (unsigned long long*)(c->dArgs + offProg);
        SearchCtl* dCtl = (SearchCtl*)(dProg + (size_t)handRows * g.wCU);      /* tickets + error word, zeroed with the hand-off rows */
        /* plain wavefront kernel: one CTA per row group, one warp per CU row */
        if (classItems[3] > classItems[2] && c->plainOct)
        {
            /* octet kernel: a warp = a band of 4 CU rows in lock step, an octet per CU (x265cu_search_oct.cuh) */
            const unsigned ni = (unsigned)(classItems[3] - classItems[2]);
            const int bands = (maxPlainRows + 3) / 4;
            const bool winVariant = c->plainWin != 0;
            /* a launch that fills the GPU is throughput-bound: its bands keep a distance and poll rarely */
            const bool full = ni * (unsigned)bands >= (unsigned)c->octFullWarps;
            const int octSlack = full ? c->octSlack : 0;
            const unsigned octSleep = full ? (unsigned)c->octSleepFull : (unsigned)c->octSleep;
            if (c->pb == 1)
            {
                const size_t smem = oct_smem_bytes<uint8_t>(bands, g.wCU);
                if (winVariant) oct_search_kernel<uint8_t, true><<<ni, bands * 32, smem, c->stream>>>(dJobs, dPlans, dItems + classItems[2], g, dLutC, dProg, dCtl, octSlack, octSleep);
                else oct_search_kernel<uint8_t, false><<<ni, bands * 32, smem, c->stream>>>(dJobs, dPlans, dItems + classItems[2], g, dLutC, dProg, dCtl, octSlack, octSleep);
            }
            else
            {
                const size_t smem = oct_smem_bytes<uint16_t>(bands, g.wCU);
                if (winVariant) oct_search_kernel<uint16_t, true><<<ni, bands * 32, smem, c->stream>>>(dJobs, dPlans, dItems + classItems[2], g, dLutC, dProg, dCtl, octSlack, octSleep);
                else oct_search_kernel<uint16_t, false><<<ni, bands * 32, smem, c->stream>>>(dJobs, dPlans, dItems + classItems[2], g, dLutC, dProg, dCtl, octSlack, octSleep);
            }
        }
        else if (classItems[3] > classItems[2])
        {
            const unsigned ni = (unsigned)(classItems[3] - classItems[2]);
            /* a launch that fills the GPU is bound by L1 line look-ups: it stages each CU's window in shared memory */
            const bool winVariant = c->plainWin == 1 || (c->plainWin == -1 && ni >= 1024) || (c->plainWin == -2 && c->pb == 1);
            const size_t pbRow4 = c->pb == 1 ? 4 : 8;
            /* TMA variant: the windows come by cp.async.bulk.tensor.3d, two buffers per warp (the next CU's window is requested a
             * step ahead); needs the tensor map of the frame mirrors */
            const bool tma = c->plainTma && c->tmaMapValid && (c->plainWin != 0);
            const size_t handBytes = ((size_t)maxPlainRows * g.wCU * sizeof(unsigned long long) + 127) & ~(size_t)127;
            const size_t smem = tma ? handBytes + (size_t)maxPlainRows * 2 * (c->pb == 1 ? (size_t)PlainTma<uint8_t>::PITCH : (size_t)PlainTma<uint16_t>::PITCH) * pbRow4 + (size_t)maxPlainRows * 2 * sizeof(unsigned long long)
                                    : (size_t)maxPlainRows * g.wCU * sizeof(unsigned long long) + (winVariant ? (size_t)maxPlainRows * WIN_PITCH * pbRow4 : 0);
#define PLAIN_LAUNCH(P, W_, O_, T_) plain_search_kernel<P, W_, O_, T_><<<ni, maxPlainRows * 32, smem, c->stream>>>(dJobs, dPlans, dItems + classItems[2], g, dLutC, dProg, &dCtl->ticket[0], c->tmaMap)
            if (c->pb == 1)
            {
                if (tma) PLAIN_LAUNCH(uint8_t, true, false, true);
                else if (winVariant && c->plainOneShot) PLAIN_LAUNCH(uint8_t, true, true, false);
                else if (winVariant) PLAIN_LAUNCH(uint8_t, true, false, false);
                else PLAIN_LAUNCH(uint8_t, false, false, false);
            }
            else
            {
                if (tma) PLAIN_LAUNCH(uint16_t, true, false, true);
                else if (winVariant && c->plainOneShot) PLAIN_LAUNCH(uint16_t, true, true, false);
                else if (winVariant) PLAIN_LAUNCH(uint16_t, true, false, false);
                else PLAIN_LAUNCH(uint16_t, false, false, false);
            }
#undef PLAIN_LAUNCH
        }
        /* speculative path, per wave -- refine: every CU of every search in parallel (estimates + memo); then the commit
         * wavefront: one CTA per row group, one warp per CU row of the group */
        int warps = maxItemRows;
        for (int w = 0; w < 2; w++)
        {
            const unsigned np = (unsigned)(classPlans[w + 1] - classPlans[w]), ni = (unsigned)(classItems[w + 1] - classItems[w]);
            if (!np) continue;
            dim3 sgrid((unsigned)((g.nCU + SPEC_WARPS - 1) / SPEC_WARPS), np);
            if (!c->searchSpec) CU_TRY(c, cudaMemsetAsync(c->dMemo + classPlans[w] * memoPerSearch, 0xff, np * memoPerSearch, c->stream));
            /* refine iterations: parallel work that shortens the commit chains; a wave with many searches is
             * throughput-bound, not chain-bound, and gets a single one */
            const int iters = np >= 32 ? (c->searchSpec < 1 ? c->searchSpec : 1) : c->searchSpec;
            if (c->pb == 1)
            {
                for (int it = 0; it < iters; it++) refine_kernel<uint8_t><<<sgrid, SPEC_WARPS * 32, 0, c->stream>>>(dJobs, dPlans + classPlans[w], g, dLutC, it);
                search_kernel<uint8_t><<<ni, warps * 32, search_smem_bytes<uint8_t>(warps, g.wCU), c->stream>>>(dJobs, dPlans, dItems + classItems[w], g, dLutC, dProg, warps, iters > 0 ? (iters - 1) % 3 : -1, &dCtl->ticket[1 + w]);
            }
            else
            {
                for (int it = 0; it < iters; it++) refine_kernel<uint16_t><<<sgrid, SPEC_WARPS * 32, 0, c->stream>>>(dJobs, dPlans + classPlans[w], g, dLutC, it);
                search_kernel<uint16_t><<<ni, warps * 32, search_smem_bytes<uint16_t>(warps, g.wCU), c->stream>>>(dJobs, dPlans, dItems + classItems[w], g, dLutC, dProg, warps, iters > 0 ? (iters - 1) % 3 : -1, &dCtl->ticket[1 + w]);
            }
        }
        CU_TRY(c, cudaGetLastError());
        /* the control words come back with the results: a hand-off wait that gave up is reported, never silent */
        CU_TRY(c, cudaMemcpyAsync(c->hArgs + offProg, dCtl, sizeof(SearchCtl), cudaMemcpyDeviceToHost, c->stream));
    }
    if (!costIdx.empty())
    {
        KernelScope ks(c, X265CU_K_COST);
        dim3 grid(g.hCU, (unsigned)costIdx.size());
        if (c->pb == 1)
            cost_kernel<uint8_t><<<grid, 128, 0, c->stream>>>((const JobDev*)(c->dArgs + offJobs), (const int*)(c->dArgs + offCost), g);
        else
            cost_kernel<uint16_t><<<grid, 128, 0, c->stream>>>((const JobDev*)(c->dArgs + offJobs), (const int*)(c->dArgs + offCost), g);
        CU_TRY(c, cudaGetLastError());
    }
    /* the bus is idle while these kernels run: now the held-back plane copy-backs of the last pre-lookahead list go */
    if (!c->deferredPlanes.empty() && flushDeferredPlanes(c)) return X265CU_ECUDA;
    if (!devCopies.empty())
    {
        KernelScope ks(c, X265CU_K_RESULTS);
        scatter_results_kernel<<<(unsigned)devCopies.size(), 256, 0, c->stream>>>((const ScatterDev*)(c->dArgs + offScat));
        CU_TRY(c, cudaGetLastError());
    }
    const size_t back = !hostCopies.empty() ? total : (size_t)n * 32;
    CU_TRY(c, cudaMemcpyAsync(c->hStage, c->dStage, back, cudaMemcpyDeviceToHost, c->stream));
    c->stats.d2hBytes += (int64_t)(back + devCopyBytes);
    const std::chrono::steady_clock::time_point tH2 = std::chrono::steady_clock::now();
    int r = syncStream(c);
    if (r) return r;
    if (!items.empty() && ((const SearchCtl*)(c->hArgs + offProg))->error)
        return fail(c, X265CU_ECUDA, "x265cu_estimate_batch: a wavefront hand-off wait timed out (search results incomplete)");
    const std::chrono::steady_clock::time_point tH3 = std::chrono::steady_clock::now();
    {
        static const bool dbgBatch = getenv("X265CU_BATCH_DEBUG") != NULL;
        if (dbgBatch)
            fprintf(stderr, "  estimate_batch: %d jobs, %zu searches (%zu row-group items), %zu cost-only, %zu weighted; planning %.3f ms, enqueue %.3f ms, wait %.3f ms\n",
                    n, plans.size(), items.size(), costIdx.size(), weightedJobs.size(), std::chrono::duration<double, std::milli>(tH1 - tH0).count(),
                    std::chrono::duration<double, std::milli>(tH2 - tH1).count(), std::chrono::duration<double, std::milli>(tH3 - tH2).count());
    }
#ifdef X265CU_SEARCH_STATS
    if (!plans.empty() && getenv("X265CU_TRACE_BATCH"))
    {
        static int batchNo = 0;
        if (batchNo++ == atoi(getenv("X265CU_TRACE_BATCH")))
        {
            static unsigned long long tt[256][512]; static unsigned int te[256][512]; static int tn[256];
            cudaMemcpyFromSymbol(tt, g_traceT, sizeof(tt)); cudaMemcpyFromSymbol(te, g_traceE, sizeof(te)); cudaMemcpyFromSymbol(tn, g_traceN, sizeof(tn));
            unsigned long long t0 = ~0ull;
            for (int r = 0; r < (int)hCU; r++) if (tn[r] > 0 && tt[r][0] < t0) t0 = tt[r][0];
            for (int r = (int)hCU - 1; r >= 0; r--)
            {
                fprintf(stderr, "TRACE row %d:", r);
                for (int i = 0; i < tn[r]; i++) fprintf(stderr, " %llu:%u/%u/%u", (tt[r][i] - t0) / 100, te[r][i] & 0xffff, (te[r][i] >> 16) & 0xff, te[r][i] >> 24);
                fprintf(stderr, "\n");
            }
        }
    }
    if (!plans.empty() && getenv("X265CU_STATS_PER_BATCH"))
    {
        unsigned long long h[32], z[32] = { 0 };
        cudaMemcpyFromSymbol(h, g_searchStats, sizeof(h));
        cudaMemcpyToSymbol(g_searchStats, z, sizeof(z));
        fprintf(stderr, "batch: jobs %d searches %zu | cus %llu fast %llu cand-pass %llu chain-search %llu steps %llu | serial %llu: cyc wait %.0f memo %.0f cand %.0f (per pass %.0f) mvp %.0f search %.0f (per search %.0f)\n",
                n, plans.size(), h[0], h[15], h[14], h[1], h[16], h[17], (double)h[20] / (h[17] ? h[17] : 1), (double)h[21] / (h[17] ? h[17] : 1),
                (double)h[22] / (h[17] ? h[17] : 1), (double)h[25] / (h[26] ? h[26] : 1), (double)h[23] / (h[17] ? h[17] : 1), (double)h[24] / (h[17] ? h[17] : 1), (double)h[24] / (h[1] ? h[1] : 1));
    }
#endif

    /* scatter to the caller's Lowres arrays; a big batch's few hundred MB are spread over several host threads
     * (a single memcpy stream runs at a fraction of the PCIe rate the data arrived with) */
    struct Copy { void* dst; const void* src; size_t bytes; };
    std::vector<Copy> copies;
    size_t copyBytes = 0;
    for (int i = 0; i < n; i++)
    {
        const x265cu_job& j = jobs[i];
        const unsigned long long* sums = (const unsigned long long*)(c->hStage + (size_t)i * 32);
        x265cu_job_result& res = results[i];
        res.costEstRaw = (int64_t)sums[0];
        res.costEstAq = (int64_t)sums[1];
        res.intraMbs = (int32_t)sums[2];
        res.reserved = 0;
        res.costEst = j.d1 > 0 ? res.costEstRaw * 100 / (130 + c->cfg.bFrameBias) : res.costEstRaw;   /* slicetype.cpp:2053-2057 */
    }
    for (size_t k = 0; k < hostCopies.size(); k++)
    {
        Copy cp = { hostCopies[k].dst, c->hStage + hostCopies[k].srcOff, hostCopies[k].bytes };
        copies.push_back(cp); copyBytes += cp.bytes;
    }
    unsigned nThreads = 1;
    if (copyBytes > ((size_t)8 << 20))
    {
        nThreads = hostThreads();
        if (nThreads > 8) nThreads = 8;
        if (nThreads < 1) nThreads = 1;
    }
    if (copies.empty()) { }
    else if (nThreads <= 1)
        for (size_t k = 0; k < copies.size(); k++) memcpy(copies[k].dst, copies[k].src, copies[k].bytes);
    else
    {
        const Copy* cp = &copies[0];
        const size_t nc = copies.size();
        std::vector<std::thread> workers;
        for (unsigned t = 0; t < nThreads; t++)
            workers.push_back(std::thread([cp, nc, t, nThreads]() { for (size_t k = t; k < nc; k += nThreads) memcpy(cp[k].dst, cp[k].src, cp[k].bytes); }));
        for (size_t t = 0; t < workers.size(); t++) workers[t].join();
    }
    c->planMs[0] += std::chrono::duration<double, std::milli>(tP1 - tH0).count();
    c->planMs[1] += std::chrono::duration<double, std::milli>(tP2 - tP1).count();
    c->planMs[2] += std::chrono::duration<double, std::milli>(tP3 - tP2).count();
    c->planMs[3] += std::chrono::duration<double, std::milli>(tH1 - tP3).count();
    c->hostMs[0] += std::chrono::duration<double, std::milli>(tH1 - tH0).count();
    c->hostMs[1] += std::chrono::duration<double, std::milli>(tH2 - tH1).count();
    c->hostMs[2] += std::chrono::duration<double, std::milli>(tH3 - tH2).count();
    c->hostMs[3] += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tH3).count();
    c->hostCalls++;
    return X265CU_OK;
}

/* -------------------------------------------------------------------------------------------- */
int x265cu_pixelcmp_batch(x265cu_ctx* c, int kind, const void* bufA, size_t samplesA, intptr_t strideA,
                          const void* bufB, size_t samplesB, intptr_t strideB,
                          int n, const int64_t* offA, const int64_t* offB, int32_t* out)
{
    if (!c || kind < 0 || kind > 3 || !bufA || !bufB || n < 0 || (n && (!offA || !offB || !out)))
        return c ? fail(c, X265CU_EINVAL, "x265cu_pixelcmp_batch: bad argument") : X265CU_EINVAL;
    if (!n) return X265CU_OK;
    {
        /* every block must lie inside the buffers it is taken from */
        const int64_t ext = kind == 3 ? 16 : 8;
        if (strideA < ext || strideB < ext) return fail(c, X265CU_EINVAL, "x265cu_pixelcmp_batch: stride smaller than the block");
        for (int i = 0; i < n; i++)
            if (offA[i] < 0 || offB[i] < 0 || (uint64_t)(offA[i] + (ext - 1) * (int64_t)strideA + ext) > samplesA ||
                (uint64_t)(offB[i] + (ext - 1) * (int64_t)strideB + ext) > samplesB)
                return fail(c, X265CU_EINVAL, "x265cu_pixelcmp_batch: a block lies outside its buffer");
    }
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    size_t bytesA = alignUp(samplesA * c->pb + 16, 256), bytesB = alignUp(samplesB * c->pb + 16, 256);
    size_t offs = alignUp((size_t)n * 8, 256);
    size_t need = bytesA + bytesB + 2 * offs + alignUp((size_t)n * 4, 256);
    if (growDevice(c, &c->dGeneric, &c->dGenericCap, need)) return X265CU_ECUDA;
    uint8_t* dA = c->dGeneric; uint8_t* dB = dA + bytesA;
    int64_t* dOffA = (int64_t*)(dB + bytesB); int64_t* dOffB = (int64_t*)((uint8_t*)dOffA + offs);
    int* dOut = (int*)((uint8_t*)dOffB + offs);
    CU_TRY(c, cudaMemsetAsync(dA, 0, bytesA + bytesB, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dA, bufA, samplesA * c->pb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dB, bufB, samplesB * c->pb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dOffA, offA, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dOffB, offB, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    {
        KernelScope ks(c, X265CU_K_PIXEL);
        int blocks = (n * 4 + 255) / 256;
        if (c->pb == 1)
            pixelcmp_batch_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>(kind, (const uint8_t*)dA, strideA, (const uint8_t*)dB, strideB, n, dOffA, dOffB, dOut);
        else
            pixelcmp_batch_kernel<uint16_t><<<blocks, 256, 0, c->stream>>>(kind, (const uint16_t*)dA, strideA, (const uint16_t*)dB, strideB, n, dOffA, dOffB, dOut);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, dOut, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    return syncStream(c);
}

int x265cu_pixelcmp_pu(x265cu_ctx* c, int kind, int width, int height, const void* bufA, size_t samplesA, intptr_t strideA,
                       const void* bufB, size_t samplesB, intptr_t strideB, int n, const int64_t* offA, const int64_t* offB, int32_t* out)
{
    if (!c || kind < 0 || kind > 1 || !bufA || !bufB || n < 0 || (n && (!offA || !offB || !out)))
        return c ? fail(c, X265CU_EINVAL, "x265cu_pixelcmp_pu: bad argument") : X265CU_EINVAL;
    {
        /* the 25 luma PU shapes of enum LumaPU (common/primitives.h:49-61) */
        static const unsigned char shapes[25][2] = { { 4, 4 }, { 8, 8 }, { 16, 16 }, { 32, 32 }, { 64, 64 }, { 8, 4 }, { 4, 8 }, { 16, 8 }, { 8, 16 }, { 32, 16 },
            { 16, 32 }, { 64, 32 }, { 32, 64 }, { 16, 12 }, { 12, 16 }, { 16, 4 }, { 4, 16 }, { 32, 24 }, { 24, 32 }, { 32, 8 }, { 8, 32 }, { 64, 48 }, { 48, 64 },
            { 64, 16 }, { 16, 64 } };
        bool known = false;
        for (int i = 0; i < 25; i++) known = known || (shapes[i][0] == width && shapes[i][1] == height);
        if (!known) return fail(c, X265CU_EINVAL, "x265cu_pixelcmp_pu: not a luma PU shape");
    }
    if (!n) return X265CU_OK;
    if (strideA < width || strideB < width) return fail(c, X265CU_EINVAL, "x265cu_pixelcmp_pu: stride smaller than the block");
    for (int i = 0; i < n; i++)
        if (offA[i] < 0 || offB[i] < 0 || (uint64_t)(offA[i] + (int64_t)(height - 1) * strideA + width) > samplesA ||
            (uint64_t)(offB[i] + (int64_t)(height - 1) * strideB + width) > samplesB)
            return fail(c, X265CU_EINVAL, "x265cu_pixelcmp_pu: a block lies outside its buffer");
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    size_t bytesA = alignUp(samplesA * c->pb + 16, 256), bytesB = alignUp(samplesB * c->pb + 16, 256);
    size_t offs = alignUp((size_t)n * 8, 256);
    size_t need = bytesA + bytesB + 2 * offs + alignUp((size_t)n * 4, 256);
    if (growDevice(c, &c->dGeneric, &c->dGenericCap, need)) return X265CU_ECUDA;
    uint8_t* dA = c->dGeneric; uint8_t* dB = dA + bytesA;
    int64_t* dOffA = (int64_t*)(dB + bytesB); int64_t* dOffB = (int64_t*)((uint8_t*)dOffA + offs);
    int* dOut = (int*)((uint8_t*)dOffB + offs);
    CU_TRY(c, cudaMemsetAsync(dA, 0, bytesA + bytesB, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dA, bufA, samplesA * c->pb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dB, bufB, samplesB * c->pb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dOffA, offA, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dOffB, offB, (size_t)n * 8, cudaMemcpyHostToDevice, c->stream));
    {
        KernelScope ks(c, X265CU_K_PIXEL);
        const int blocks = (n + 7) / 8;
        if (c->pb == 1)
            pixelcmp_pu_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>(kind, width, height, (const uint8_t*)dA, strideA, (const uint8_t*)dB, strideB, n, dOffA, dOffB, dOut);
        else
            pixelcmp_pu_kernel<uint16_t><<<blocks, 256, 0, c->stream>>>(kind, width, height, (const uint16_t*)dA, strideA, (const uint16_t*)dB, strideB, n, dOffA, dOffB, dOut);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(out, dOut, (size_t)n * 4, cudaMemcpyDeviceToHost, c->stream));
    return syncStream(c);
}

/* full-resolution PU motion search (x265cu_me.cuh): n searches of one PU shape, a warp each */
int x265cu_motion_estimate(x265cu_ctx* c, int searchMethod, int subpelRefine, int width, int height,
                           const void* fencPlane, size_t fencSamples, intptr_t fencStride,
                           const void* refPlane, size_t refSamples, intptr_t refStride, const uint16_t* mvcostCentre,
                           int n, const x265cu_me_item* items, x265cu_me_result* out, float* ms)
{
    if (!c || searchMethod < 0 || searchMethod > 4 || subpelRefine < 0 || subpelRefine > 7 || !fencPlane || !refPlane || !mvcostCentre ||
        n < 0 || (n && (!items || !out)))
        return c ? fail(c, X265CU_EINVAL, "x265cu_motion_estimate: bad argument") : X265CU_EINVAL;
    {
        static const unsigned char shapes[24][2] = { { 8, 8 }, { 16, 16 }, { 32, 32 }, { 64, 64 }, { 8, 4 }, { 4, 8 }, { 16, 8 }, { 8, 16 }, { 32, 16 },
            { 16, 32 }, { 64, 32 }, { 32, 64 }, { 16, 12 }, { 12, 16 }, { 16, 4 }, { 4, 16 }, { 32, 24 }, { 24, 32 }, { 32, 8 }, { 8, 32 }, { 64, 48 }, { 48, 64 },
            { 64, 16 }, { 16, 64 } };
        bool known = false;
        for (int i = 0; i < 24; i++) known = known || (shapes[i][0] == width && shapes[i][1] == height);
        if (!known) return fail(c, X265CU_EINVAL, "x265cu_motion_estimate: not an inter PU shape (4x4 is not one: motion.cpp:168)");
    }
    if (ms) *ms = 0.f;
    if (!n) return X265CU_OK;
    if (fencStride < width || refStride < width) return fail(c, X265CU_EINVAL, "x265cu_motion_estimate: stride smaller than the block");
    for (int i = 0; i < n; i++)
    {
        const x265cu_me_item& it = items[i];
        if (it.numCandidates < 0 || it.numCandidates > 12 || it.merange < 1 || it.merange > 1024 || it.mvmin[0] > it.mvmax[0] || it.mvmin[1] > it.mvmax[1] ||
            it.mvmin[0] < -8000 || it.mvmin[1] < -8000 || it.mvmax[0] > 8000 || it.mvmax[1] > 8000)
            return fail(c, X265CU_EINVAL, "x265cu_motion_estimate: bad item (candidates 0..12, merange 1..1024, mvmin <= mvmax, |mv| <= 8000)");
        const int64_t lo = it.offset + (int64_t)(it.mvmin[1] - 16) * refStride + it.mvmin[0] - 16;
        const int64_t hi = it.offset + (int64_t)(it.mvmax[1] + 16 + height - 1) * refStride + it.mvmax[0] + 16 + width;
        if (it.offset < 0 || lo < 0 || (uint64_t)hi > refSamples || (uint64_t)(it.offset + (int64_t)(height - 1) * fencStride + width) > fencSamples)
            return fail(c, X265CU_EINVAL, "x265cu_motion_estimate: a search window (+16 samples) leaves the reference plane, or a PU its source plane");
    }
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const size_t bytesA = alignUp(fencSamples * c->pb + 16, 256), bytesB = alignUp(refSamples * c->pb + 16, 256);
    const size_t lutBytes = alignUp((size_t)131073 * 2, 256), itemBytes = alignUp((size_t)n * sizeof(x265cu_me_item), 256);
    const size_t need = bytesA + bytesB + lutBytes + itemBytes + alignUp((size_t)n * sizeof(x265cu_me_result), 256);
    if (growDevice(c, &c->dGeneric, &c->dGenericCap, need)) return X265CU_ECUDA;
    uint8_t* dA = c->dGeneric; uint8_t* dB = dA + bytesA;
    uint16_t* dLut = (uint16_t*)(dB + bytesB);
    MeItemDev* dItems = (MeItemDev*)((uint8_t*)dLut + lutBytes);
    MeResultDev* dOut = (MeResultDev*)((uint8_t*)dItems + itemBytes);
    static_assert(sizeof(MeItemDev) == sizeof(x265cu_me_item) && sizeof(MeResultDev) == sizeof(x265cu_me_result), "item layout");
    CU_TRY(c, cudaMemcpyAsync(dA, fencPlane, fencSamples * c->pb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dB, refPlane, refSamples * c->pb, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemsetAsync(dB + refSamples * c->pb, 0, bytesB - refSamples * c->pb, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dLut, mvcostCentre - 65536, (size_t)131073 * 2, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(c, cudaMemcpyAsync(dItems, items, (size_t)n * sizeof(x265cu_me_item), cudaMemcpyHostToDevice, c->stream));
    const int warps = 4;
    const size_t blockBytes = (size_t)width * height * c->pb;
    const size_t perWarp = 2 * blockBytes + alignUp((size_t)(height + 7) * width * 2, 16) + 16 * sizeof(MePt) + 16 * sizeof(int);
    const size_t smem = warps * perWarp;
    cudaEvent_t e0 = getEvent(c), e1 = getEvent(c);
    c->stats.launches[X265CU_K_PIXEL]++;
    const int blocks = (n + warps - 1) / warps;
    if (c->pb == 1)
    {
        CU_TRY(c, cudaFuncSetAttribute(pu_motion_search_kernel<uint8_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(c, cudaEventRecord(e0, c->stream));
        pu_motion_search_kernel<uint8_t><<<blocks, warps * 32, smem, c->stream>>>(searchMethod, subpelRefine, width, height, c->cfg.bitDepth, (const uint8_t*)dA, fencStride,
                                                                                 (const uint8_t*)dB, refStride, dLut + 65536, n, dItems, dOut);
    }
    else
    {
        CU_TRY(c, cudaFuncSetAttribute(pu_motion_search_kernel<uint16_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CU_TRY(c, cudaEventRecord(e0, c->stream));
        pu_motion_search_kernel<uint16_t><<<blocks, warps * 32, smem, c->stream>>>(searchMethod, subpelRefine, width, height, c->cfg.bitDepth, (const uint16_t*)dA, fencStride,
                                                                                  (const uint16_t*)dB, refStride, dLut + 65536, n, dItems, dOut);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaEventRecord(e1, c->stream));
    CU_TRY(c, cudaMemcpyAsync(out, dOut, (size_t)n * sizeof(x265cu_me_result), cudaMemcpyDeviceToHost, c->stream));
    int rc = syncStream(c);
    if (rc) return rc;
    if (ms) CU_TRY(c, cudaEventElapsedTime(ms, e0, e1));
    return X265CU_OK;
}

int x265cu_pixelcmp_frames(x265cu_ctx* c, int kind, int nPairs, const int* slotsA, const int* slotsB, int32_t* out, float* ms)
{
    return x265cu_pixelcmp_planes(c, kind, nPairs, slotsA, NULL, slotsB, NULL, out, ms);
}

int x265cu_pixelcmp_planes(x265cu_ctx* c, int kind, int nPairs, const int* slotsA, const int* planesA, const int* slotsB, const int* planesB,
                           int32_t* out, float* ms)
{
    if (!c || kind < 0 || kind > 2 || nPairs < 1 || nPairs > 4096 || !slotsA || !slotsB)
        return c ? fail(c, X265CU_EINVAL, "x265cu_pixelcmp_frames: bad argument") : X265CU_EINVAL;
    for (int i = 0; i < nPairs; i++)
        if (badSlot(c, slotsA[i]) || badSlot(c, slotsB[i]) || (planesA && (planesA[i] < 0 || planesA[i] > 3)) || (planesB && (planesB[i] < 0 || planesB[i] > 3)))
            return fail(c, X265CU_EINVAL, "x265cu_pixelcmp_frames: bad slot or plane");
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const GeomDev& g = c->g;
    const size_t outBytes = alignUp((size_t)nPairs * g.nCU * 4, 256);
    if (growDevice(c, &c->dGeneric, &c->dGenericCap, outBytes + (size_t)nPairs * 16)) return X265CU_ECUDA;
    const void** hp = (const void**)malloc((size_t)nPairs * 16);
    const size_t planeBytes = (size_t)c->g.planeSize * c->pb;
    for (int i = 0; i < nPairs; i++)
    {
        hp[2 * i] = slotPlane0(c, slotsA[i]) + (planesA ? planesA[i] : 0) * planeBytes;
        hp[2 * i + 1] = slotPlane0(c, slotsB[i]) + (planesB ? planesB[i] : 0) * planeBytes;
    }
    cudaError_t ce = cudaMemcpyAsync(c->dGeneric + outBytes, hp, (size_t)nPairs * 16, cudaMemcpyHostToDevice, c->stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(c->stream);
    free(hp);
    CU_TRY(c, ce);
    cudaEvent_t e0 = getEvent(c), e1 = getEvent(c);
    c->stats.launches[X265CU_K_PIXEL]++;
    CU_TRY(c, cudaEventRecord(e0, c->stream));
    /* SAD / SATD: the wide form (X265CU_PIXELCMP_WIDE: 1 = one group of 16 CUs per warp iteration (default; measured 0.91 of
     * the HBM roofline at 8 bit against 0.65 for the quad form), 2 = two groups, 0 = the quad form, which SA8D always takes) */
    int wide = kind < 2 ? 1 : 0;
    if (const char* e = getenv("X265CU_PIXELCMP_WIDE")) { int v = atoi(e); if (kind < 2 && v >= 0 && v <= 2) wide = v; }
    /* its 8-sample loads need planes whose rows start on 8-sample boundaries (true for every x265 margin: CTU + 32) */
    if ((g.padOffset & 7) || (g.stride & 7) || (g.planeSize & 7)) wide = 0;
    const int cusPerWarp = wide ? 16 * wide : 8;
    int bx = ((g.nCU + cusPerWarp - 1) / cusPerWarp + 7) / 8;
    if (bx > 148 * 4) bx = 148 * 4;
    /* wide form, 8-bit samples: at most 32 CTAs per pair, their warps loop over the rest of the plane (measured cold at 4K,
     * fraction of the HBM roofline: 254 CTAs per pair 0.75, 32: 0.82, 16: 0.80, 8: 0.73, 4: 0.59; at 16 bit, where a warp
     * moves twice the bytes, the full grid measured better: 0.94 against 0.89 -- profiles/README.md) */
    if (wide && c->pb == 1 && bx > 32) bx = 32;
    if (const char* e = getenv("X265CU_PIXELCMP_BX")) { int v = atoi(e); if (v >= 1 && v < bx) bx = v; }   /* experiments */
    if (bx < 1) bx = 1;
    dim3 grid(bx, nPairs);
    const void* const* dPl = (const void* const*)(c->dGeneric + outBytes);
    if (c->pb == 1)
    {
        if (wide == 2) pixelcmp_frames_wide_kernel<uint8_t, 2><<<grid, 256, 0, c->stream>>>(kind, dPl, g, (int*)c->dGeneric);
        else if (wide == 1) pixelcmp_frames_wide_kernel<uint8_t, 1><<<grid, 256, 0, c->stream>>>(kind, dPl, g, (int*)c->dGeneric);
        else pixelcmp_frames_kernel<uint8_t><<<grid, 256, 0, c->stream>>>(kind, dPl, g, (int*)c->dGeneric);
    }
    else
    {
        if (wide == 2) pixelcmp_frames_wide_kernel<uint16_t, 2><<<grid, 256, 0, c->stream>>>(kind, dPl, g, (int*)c->dGeneric);
        else if (wide == 1) pixelcmp_frames_wide_kernel<uint16_t, 1><<<grid, 256, 0, c->stream>>>(kind, dPl, g, (int*)c->dGeneric);
        else pixelcmp_frames_kernel<uint16_t><<<grid, 256, 0, c->stream>>>(kind, dPl, g, (int*)c->dGeneric);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaEventRecord(e1, c->stream));
    if (out) CU_TRY(c, cudaMemcpyAsync(out, c->dGeneric, (size_t)nPairs * g.nCU * 4, cudaMemcpyDeviceToHost, c->stream));
    int r = syncStream(c);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    c->stats.ms[X265CU_K_PIXEL] += t;
    if (ms) *ms = t;
    c->freeEvents.push_back(e0); c->freeEvents.push_back(e1);
    return r;
}

/* integer-pipe micro-benchmark: the roofline of the search/cost kernels is integer issue rate, not
 * HBM, and MEASURED_PEAKS.json has no integer peak.  Times a dependent-free stream of packed
 * absolute-difference-accumulate (the SAD inner op) and of IADD3-class adds on every SM. */
int x265cu_int_peak(x265cu_ctx* c, double* gopsVabsdiff4, double* gopsIadd)
{
    if (!c) return X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    if (growDevice(c, &c->dGeneric, &c->dGenericCap, 1 << 20)) return X265CU_ECUDA;
    const int iters = 4096, blocks = 148 * 8, threads = 256;
    double res[2] = { 0, 0 };
    for (int mode = 0; mode < 2; mode++)
    {
        float best = 1e30f;
        for (int rep = 0; rep < 4; rep++)
        {
            cudaEvent_t e0 = getEvent(c), e1 = getEvent(c);
            CU_TRY(c, cudaEventRecord(e0, c->stream));
            int_peak_kernel<<<blocks, threads, 0, c->stream>>>(mode, iters, (unsigned int*)c->dGeneric);
            CU_TRY(c, cudaEventRecord(e1, c->stream));
            CU_TRY(c, cudaStreamSynchronize(c->stream));
            float t = 0;
            cudaEventElapsedTime(&t, e0, e1);
            if (rep && t < best) best = t;
            c->freeEvents.push_back(e0); c->freeEvents.push_back(e1);
        }
        /* 16 independent ops per iteration per thread */
        res[mode] = (double)blocks * threads * iters * 16 / (best * 1e-3) / 1e9;
    }
    CU_TRY(c, cudaGetLastError());
    if (gopsVabsdiff4) *gopsVabsdiff4 = res[0];
    if (gopsIadd) *gopsIadd = res[1];
    return X265CU_OK;
}

int x265cu_frame_var(x265cu_ctx* c, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride,
                     int planesAreDevice, uint32_t* energy, uint64_t sums[6])
{
    if (!c || !y || !energy || !sums || ((u == NULL) != (v == NULL)))
        return c ? fail(c, X265CU_EINVAL, "x265cu_frame_var: bad argument") : X265CU_EINVAL;
    std::lock_guard<std::mutex> lk(c->mtx);
    CU_TRY(c, cudaSetDevice(c->cfg.device));
    const int W = c->cfg.srcWidth, H = c->cfg.srcHeight;
    const int bxN = (W + 15) / 16, byN = (H + 15) / 16;
    const size_t yp = alignUp((size_t)bxN * 16, 64), cp = alignUp((size_t)bxN * 8, 64);
    const size_t yBytes = yp * byN * 16 * c->pb, cBytes = cp * byN * 8 * c->pb;
    const size_t eBytes = alignUp((size_t)bxN * byN * 4, 256);
    if (growDevice(c, &c->dGeneric, &c->dGenericCap, yBytes + 2 * cBytes + eBytes + 256)) return X265CU_ECUDA;
    uint8_t* dY = c->dGeneric; uint8_t* dU = dY + yBytes; uint8_t* dV = dU + cBytes; unsigned int* dE = (unsigned int*)(dV + cBytes);
    int64_t ypitch = (int64_t)yp, cpitch = (int64_t)cp;
    if (planesAreDevice)
    {
        dY = (uint8_t*)y; dU = (uint8_t*)u; dV = (uint8_t*)v;
        ypitch = yStride; cpitch = cStride;
    }
    else
    {
        CU_TRY(c, cudaMemcpy2DAsync(dY, yp * c->pb, y, (size_t)yStride * c->pb, (size_t)bxN * 16 * c->pb, byN * 16, cudaMemcpyHostToDevice, c->stream));
        if (u)
        {
            CU_TRY(c, cudaMemcpy2DAsync(dU, cp * c->pb, u, (size_t)cStride * c->pb, (size_t)bxN * 8 * c->pb, byN * 8, cudaMemcpyHostToDevice, c->stream));
            CU_TRY(c, cudaMemcpy2DAsync(dV, cp * c->pb, v, (size_t)cStride * c->pb, (size_t)bxN * 8 * c->pb, byN * 8, cudaMemcpyHostToDevice, c->stream));
        }
        c->stats.h2dBytes += (int64_t)(yBytes + (u ? 2 * cBytes : 0));
    }
    CU_TRY(c, cudaMemsetAsync(c->dSmall, 0, 6 * sizeof(unsigned long long), c->stream));
    {
        KernelScope ks(c, X265CU_K_VAR);
        int blocks = (bxN * byN + 7) / 8;
            if (blocks > 148 * 8) blocks = 148 * 8;      /* warps stride over the 16x16 blocks */
        if (c->pb == 1)
            frame_var_kernel<uint8_t><<<blocks, 256, 0, c->stream>>>((const uint8_t*)dY, ypitch, u ? (const uint8_t*)dU : NULL, u ? (const uint8_t*)dV : NULL, cpitch, bxN, byN, dE, c->dSmall, NULL, NULL);
        else
            frame_var_kernel<uint16_t><<<blocks, 256, 0, c->stream>>>((const uint16_t*)dY, ypitch, u ? (const uint16_t*)dU : NULL, u ? (const uint16_t*)dV : NULL, cpitch, bxN, byN, dE, c->dSmall, NULL, NULL);
    }
    CU_TRY(c, cudaGetLastError());
    CU_TRY(c, cudaMemcpyAsync(energy, dE, (size_t)bxN * byN * 4, cudaMemcpyDeviceToHost, c->stream));
    unsigned long long hs[6];
    CU_TRY(c, cudaMemcpyAsync(hs, c->dSmall, sizeof(hs), cudaMemcpyDeviceToHost, c->stream));
    int r = syncStream(c);
    for (int i = 0; i < 6; i++) sums[i] = hs[i];
    return r;
}

} // extern "C"
