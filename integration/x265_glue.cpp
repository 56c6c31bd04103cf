/* x265_glue.cpp -- the x265-side binding of libx265cu.so, inside the real x265 1.9 encoder (see x265_glue.h).
 *
 * Compiled WITH the reference's headers and linked with the reference's objects, libx265cu_host.so and libx265cu.so
 * (integration/build_x265_cu.py).  It is a thin adapter: every x265 `Lowres` gets a shadow `x265cu::Lowres` whose ARRAYS
 * ARE x265's own arrays (pinned + mapped, so the GPU writes results in place), the call-outs translate x265's calls into
 * the host layer's (src/x265_b200/host/lookahead_cu.{h,cpp}: batching, look-ahead estimate cache, weightsAnalyse float
 * guesses, calcAdaptiveQuantFrame float mapping, cuTree queue) and mirror the scalar members (costEst, costEstAq, intraMbs,
 * weightedCostDelta, wp_sum/wp_ssd) back into x265's struct.  Everything else -- slicetypeDecide, slicetypeAnalyse,
 * scenecut, slicetypePath, cuTree's control flow and cuTreeFinish's log2 mapping, vbvLookahead, frameCostRecalculate, rate
 * control -- is x265's unchanged code consuming those arrays.  No CPU fallback: a failing GPU call aborts the encoder.
 */
#include "common.h"
#include "frame.h"
#include "picyuv.h"
#include "lowres.h"
#include "slicetype.h"
#include "threadpool.h"

#include "x265_glue.h"
#include "x265cu.h"
#include "lookahead_cu.h"

#include <map>
#include <vector>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>

namespace xr = X265_NS;

/* observation hooks of the trace harness (harness/x265_la_driver.cpp); absent in the CLI build */
extern "C" {
void x265ref_hook_pre(xr::Frame* frame) __attribute__((weak));
void x265ref_hook_batch(int begin, int njobs) __attribute__((weak));
void x265ref_hook_job(xr::Lowres** frames, int p0, int p1, int b, int search0, int search1, int batchMode, int sliced) __attribute__((weak));
void x265ref_hook_weight(int fencPoc, int refPoc, int scale, int denom, int offset) __attribute__((weak));
void x265ref_hook_ctzero(xr::Lowres* frame) __attribute__((weak));
void x265ref_hook_propagate(xr::Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced) __attribute__((weak));
void x265ref_hook_ctfinish(xr::Lowres* frame, double averageDuration, int ref0Distance) __attribute__((weak));
int x265la_trace_level(void) __attribute__((weak));
}

namespace {

void die(const char* what, const char* why)
{
    fprintf(stderr, "x265 [error]: lookahead GPU path: %s: %s\n", what, why ? why : "");
    abort();      /* no CPU fallback */
}

int traceLevel() { return x265la_trace_level ? x265la_trace_level() : 0; }

/* ------------------------------------------------------------------------------------------------ pinned memory */
struct Block { size_t bytes; int live; bool arena; };
pthread_mutex_t g_memLock = PTHREAD_MUTEX_INITIALIZER;
std::map<uintptr_t, Block> g_blocks;            /* base -> block */
__thread uint8_t* t_arena;                      /* arena of the Lowres::create running on this thread */
__thread size_t t_arenaBytes, t_arenaUsed;

void* pinnedBlock(size_t bytes, bool arena)
{
    void* p = NULL;
    bytes = (bytes + 4095) & ~(size_t)4095;
    if (posix_memalign(&p, 4096, bytes)) return NULL;
    x265cu_host_register(p, bytes);             /* failure (no device here) only costs speed: the memory is still memory */
    Block b = { bytes, 0, arena };
    pthread_mutex_lock(&g_memLock);
    g_blocks[(uintptr_t)p] = b;
    pthread_mutex_unlock(&g_memLock);
    return p;
}

struct GlueState
{
    x265cu::Lookahead la;
    std::map<xr::Lowres*, x265cu::Lowres*> shadows;
    pthread_mutex_t mapLock;        /* the frame encoders look shadows up (weightAnalyse) while the lookahead adds / evicts them */
    pthread_mutex_t wpLock;         /* one weightAnalyse at a time per context (x265cu_wp_prepare .. x265cu_wp_cost) */
    /* X265CU_GLUE_PROFILE=1: wall seconds and calls per call-out, printed by x265glue_close */
    bool profile;
    double secs[6];
    long calls[6];
    struct timespec opened;
};

enum { T_PRE = 0, T_SINGLE, T_BATCH, T_CTFETCH, T_SYNC, T_OTHER };

struct Timed
{
    GlueState* st; int k; struct timespec t0;
    Timed(GlueState* s, int kind) : st(s), k(kind) { if (st->profile) clock_gettime(CLOCK_MONOTONIC, &t0); }
    ~Timed()
    {
        if (!st->profile) return;
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        st->secs[k] += (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
        st->calls[k]++;
    }
};

pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
std::map<xr::Lookahead*, GlueState*> g_states;
int64_t g_totals[8];        /* over every context closed so far: contexts, h2d bytes, d2h bytes, kernel launches, look-ahead cache stats[4] */

GlueState* stateOf(xr::Lookahead* la)
{
    pthread_mutex_lock(&g_lock);
    std::map<xr::Lookahead*, GlueState*>::iterator it = g_states.find(la);
    GlueState* st = it == g_states.end() ? NULL : it->second;
    pthread_mutex_unlock(&g_lock);
    if (!st) die("no GPU context for this Lookahead", "x265glue_open was not called (Lookahead::create)");
    return st;
}

void fillArrays(x265cu::Lowres& a, xr::Lowres* xl, int bframes)
{
    memset(&a, 0, sizeof(a));
    for (int i = 0; i < 4; i++) { a.buffer[i] = xl->buffer[i]; a.lowresPlane[i] = xl->lowresPlane[i]; }
    a.intraCost = xl->intraCost; a.intraMode = xl->intraMode; a.propagateCost = xl->propagateCost;
    a.qpAqOffset = xl->qpAqOffset; a.qpCuTreeOffset = xl->qpCuTreeOffset; a.invQscaleFactor = xl->invQscaleFactor;
    a.blockVariance = xl->blockVariance;
    for (int i = 0; i < bframes + 2; i++)
        for (int j = 0; j < bframes + 2; j++) { a.rowSatds[i][j] = xl->rowSatds[i][j]; a.lowresCosts[i][j] = xl->lowresCosts[i][j]; }
    for (int i = 0; i < bframes + 1; i++)
        for (int l = 0; l < 2; l++)
        {
            a.lowresMvs[l][i] = reinterpret_cast<x265cu::MV*>(xl->lowresMvs[l][i]);
            a.lowresMvCosts[l][i] = xl->lowresMvCosts[l][i];
        }
}

/* the shadow of an x265 Lowres (created on first sight).  Out of device slots: frames older than the last non-B frame
 * never take part in an estimate again (frames[0] = m_lastNonB is the oldest frame any call names) and give theirs up. */
struct MapGuard
{
    pthread_mutex_t* m;
    MapGuard(pthread_mutex_t* mm) : m(mm) { pthread_mutex_lock(m); }
    ~MapGuard() { pthread_mutex_unlock(m); }
};

x265cu::Lowres* shadowOf(GlueState* st, xr::Lookahead* xla, xr::Lowres* xl)
{
    MapGuard guard(&st->mapLock);
    std::map<xr::Lowres*, x265cu::Lowres*>::iterator it = st->shadows.find(xl);
    if (it != st->shadows.end())
    {
        x265cu::Lowres* sh = it->second;
        if (sh->buffer[0] == xl->buffer[0] && sh->intraCost == xl->intraCost)
        {
            sh->propagateCost = xl->propagateCost;      /* cuTree swaps these pointers (rc-lookahead 0) */
            return sh;
        }
        st->la.freeLowres(sh);                          /* the Lowres at this address was destroyed and re-created */
        st->shadows.erase(it);
    }
    if (st->la.m_freeSlots.empty())
    {
        const int oldest = xla->m_lastNonB ? xla->m_lastNonB->frameNum : 0;
        for (it = st->shadows.begin(); it != st->shadows.end();)
            if (it->first != xl && it->first != xla->m_lastNonB && it->second->frameNum < oldest)
            {
                st->la.freeLowres(it->second);
                st->shadows.erase(it++);
            }
            else
                ++it;
    }
    x265cu::Lowres a;
    fillArrays(a, xl, st->la.m_param.bframes);
    if (xl->lumaStride != st->la.m_geom.stride || xl->buffer[1] - xl->buffer[0] != st->la.m_geom.planeSize ||
        xl->lowresPlane[0] - xl->buffer[0] != st->la.m_geom.padOffset)
        die("Lowres geometry", "Lowres::create and x265cu_get_geometry disagree");
    x265cu::Lowres* sh = st->la.adoptLowres(a);
    if (!sh) die("adoptLowres", st->la.m_error);
    sh->frameNum = xl->frameNum;                        /* (Lowres::init ran: x265glue_pre_list) */
    st->shadows[xl] = sh;
    return sh;
}

x265cu::Lowres* knownShadow(GlueState* st, xr::Lookahead* xla, xr::Lowres* xl)
{
    x265cu::Lowres* sh = shadowOf(st, xla, xl);
    if (!sh->ready) die("estimate", "a frame that did not go through the pre-lookahead");
    return sh;
}

/* what estimateFrameCost leaves in the struct besides the arrays (slicetype.cpp:1999-2057) */
void mirrorEstimate(xr::Lowres* fenc, const x265cu::Lowres* sh, int d0, int d1)
{
    fenc->costEst[d0][d1] = sh->costEst[d0][d1];
    fenc->costEstAq[d0][d1] = sh->costEstAq[d0][d1];
    fenc->intraMbs[d0] = sh->intraMbs[d0];
    fenc->weightedRef[d0].isWeighted = sh->weightedRef[d0].present != 0;
    fenc->weightedCostDelta[d0] = sh->weightedCostDelta[d0];
}

void traceJob(xr::Lowres** frames, const x265cu::Lowres* sh, int p0, int p1, int b, int s0, int s1, int batchMode, int sliced)
{
    if (!traceLevel() || !x265ref_hook_job) return;
    const int d0 = b - p0;
    if (s0 && sh->weightedRef[d0].present && x265ref_hook_weight)
        x265ref_hook_weight(frames[b]->frameNum, frames[p0]->frameNum, sh->weightedRef[d0].scale, sh->weightedRef[d0].denom, sh->weightedRef[d0].offset);
    /* the reference's observation point lies before the B-frame scaling of the score (slicetype.cpp:2053) */
    const int d1 = p1 - b;
    const int64_t stored = frames[b]->costEst[d0][d1];
    frames[b]->costEst[d0][d1] = sh->costEstRaw[d0][d1];
    x265ref_hook_job(frames, p0, p1, b, s0, s1, batchMode, sliced);
    frames[b]->costEst[d0][d1] = stored;
}

} // namespace

/* ================================================================================================ memory */
extern "C" void* x265glue_malloc(size_t bytes)
{
    if (t_arena)
    {
        size_t at = (t_arenaUsed + 63) & ~(size_t)63;
        if (at + bytes <= t_arenaBytes)
        {
            t_arenaUsed = at + bytes;
            pthread_mutex_lock(&g_memLock);
            g_blocks[(uintptr_t)t_arena].live++;
            pthread_mutex_unlock(&g_memLock);
            return t_arena + at;
        }
    }
    void* p = pinnedBlock(bytes, false);
    if (p)
    {
        pthread_mutex_lock(&g_memLock);
        g_blocks[(uintptr_t)p].live = 1;
        pthread_mutex_unlock(&g_memLock);
    }
    return p;
}

extern "C" void x265glue_free(void* p)
{
    if (!p) return;
    void* release = NULL;
    pthread_mutex_lock(&g_memLock);
    std::map<uintptr_t, Block>::iterator it = g_blocks.upper_bound((uintptr_t)p);
    if (it != g_blocks.begin())
    {
        --it;
        if ((uintptr_t)p < it->first + it->second.bytes)
        {
            if (--it->second.live <= 0)
            {
                release = (void*)it->first;
                g_blocks.erase(it);
            }
            pthread_mutex_unlock(&g_memLock);
            if (release)
            {
                x265cu_host_unregister(release);
                free(release);
            }
            return;
        }
    }
    pthread_mutex_unlock(&g_memLock);
    free(p);        /* not ours */
}

extern "C" void x265glue_arena_begin(size_t bytes)
{
    t_arena = (uint8_t*)pinnedBlock(bytes, true);
    t_arenaBytes = t_arena ? ((bytes + 4095) & ~(size_t)4095) : 0;
    t_arenaUsed = 0;
}

extern "C" void x265glue_arena_end(void)
{
    if (t_arena)
    {
        /* nothing was carved out of it (Lowres::create failed early): give it back */
        void* release = NULL;
        pthread_mutex_lock(&g_memLock);
        std::map<uintptr_t, Block>::iterator it = g_blocks.find((uintptr_t)t_arena);
        if (it != g_blocks.end() && it->second.live == 0) { release = t_arena; g_blocks.erase(it); }
        pthread_mutex_unlock(&g_memLock);
        if (release) { x265cu_host_unregister(release); free(release); }
    }
    t_arena = NULL; t_arenaBytes = t_arenaUsed = 0;
}

/* upper bound of what Lowres::create allocates (lowres.cpp:30-95), each array padded to 64 bytes */
extern "C" size_t x265glue_lowres_bytes(int picWidth, int picHeight, int marginX, int marginY, int bframes, int pixelBytes)
{
    int width = picWidth / 2, lines = picHeight / 2;
    size_t stride = (size_t)width + 2 * marginX;
    if (stride & 31) stride += 32 - (stride & 31);
    const size_t cols = (width + 7) >> 3, rows = (lines + 7) >> 3, n = cols * rows;
    lines = (int)rows * 8;
    const size_t planesize = stride * (lines + 2 * marginY);
    size_t bytes = 4 * planesize * pixelBytes + 64;
    bytes += n * (8 + 4 + 8 + 4) + 4 * 64;                              /* qpAqOffset, invQscaleFactor, qpCuTreeOffset, blockVariance */
    bytes += n * (2 + 4 + 1) + 3 * 64;                                  /* propagateCost, intraCost, intraMode */
    bytes += (size_t)(bframes + 2) * (bframes + 2) * (rows * 4 + n * 2 + 2 * 64);
    bytes += (size_t)(bframes + 1) * 4 * (n * 4 + 64);
    return bytes + 4096;
}

extern "C" int x265glue_active(void) { return 1; }

/* ================================================================================================ context */
extern "C" void x265glue_open(xr::Lookahead* la)
{
    const x265_param* p = la->m_param;
    if (p->rc.bStatRead && p->rc.cuTree)
        die("x265glue_open", "cuTree offsets from a stats file (2-pass) are not supported by the GPU lookahead");
    x265cu::Param q;
    memset(&q, 0, sizeof(q));
    q.sourceWidth = p->sourceWidth; q.sourceHeight = p->sourceHeight;
    q.bitDepth = X265_DEPTH;
    q.maxCUSize = (int)xr::g_maxCUSize;
    q.bframes = p->bframes;
    q.lookaheadDepth = p->lookaheadDepth;
    q.lookaheadSlices = p->lookaheadSlices;
    q.poolWorkers = la->m_pool ? la->m_pool->m_numWorkers : 0;
    q.forceCoopSlices = la->m_numCoopSlices; q.forceRowsPerSlice = la->m_numRowsPerSlice;   /* slicetype.cpp:534-558, as computed there */
    q.bEnableWeightedPred = p->bEnableWeightedPred;
    q.bEnableWeightedBiPred = p->bEnableWeightedBiPred;
    q.bBPyramid = p->bBPyramid;
    q.aqMode = p->rc.aqMode; q.aqStrength = p->rc.aqStrength;
    q.bFrameBias = p->bFrameBias;
    q.device = getenv("X265CU_DEVICE") ? atoi(getenv("X265CU_DEVICE")) : 0;
    q.frameSlots = getenv("X265CU_FRAME_SLOTS") ? atoi(getenv("X265CU_FRAME_SLOTS"))
                                                : p->lookaheadDepth + p->bframes + 2 * X265_MAX(p->frameNumThreads, 1) + 24;
    q.fpsNum = (int)p->fpsNum; q.fpsDenom = (int)p->fpsDenom;
    q.qCompress = p->rc.qCompress;
    GlueState* st = new GlueState;
    pthread_mutex_init(&st->mapLock, NULL);
    pthread_mutex_init(&st->wpLock, NULL);
    st->profile = getenv("X265CU_GLUE_PROFILE") && atoi(getenv("X265CU_GLUE_PROFILE"));
    memset(st->secs, 0, sizeof(st->secs)); memset(st->calls, 0, sizeof(st->calls));
    if (!st->la.create(q)) die("x265cu_open", st->la.m_error);
    if (st->la.m_8x8Width != la->m_8x8Width || st->la.m_8x8Height != la->m_8x8Height || st->la.m_8x8Blocks != la->m_8x8Blocks ||
        st->la.m_numCoopSlices != la->m_numCoopSlices || st->la.m_numRowsPerSlice != la->m_numRowsPerSlice)
        die("x265glue_open", "lookahead geometry differs from Lookahead::Lookahead");
    clock_gettime(CLOCK_MONOTONIC, &st->opened);
    pthread_mutex_lock(&g_lock);
    g_states[la] = st;
    pthread_mutex_unlock(&g_lock);
}

extern "C" void x265glue_close(xr::Lookahead* la)
{
    pthread_mutex_lock(&g_lock);
    std::map<xr::Lookahead*, GlueState*>::iterator it = g_states.find(la);
    GlueState* st = it == g_states.end() ? NULL : it->second;
    if (st) g_states.erase(it);
    pthread_mutex_unlock(&g_lock);
    if (!st) return;
    {
        x265cu_stats xs;
        if (x265cu_stats_get(st->la.m_ctx, &xs, 0) == 0)
        {
            pthread_mutex_lock(&g_lock);
            g_totals[0]++; g_totals[1] += xs.h2dBytes; g_totals[2] += xs.d2hBytes;
            for (int k = 0; k < X265CU_K_COUNT; k++) g_totals[3] += xs.launches[k];
            for (int k = 0; k < 4; k++) g_totals[4 + k] += st->la.m_specStats[k];
            pthread_mutex_unlock(&g_lock);
        }
    }
    if (st->profile)
    {
        struct timespec t1;
        clock_gettime(CLOCK_MONOTONIC, &t1);
        static const char* names[6] = { "pre_list", "singleCost", "finishBatch", "cuTree fetch", "sync", "other" };
        double sum = 0;
        for (int k = 0; k < 6; k++) sum += st->secs[k];
        fprintf(stderr, "x265glue: %.1f ms since open, %.1f ms inside the GPU call-outs:", 1e3 * ((double)(t1.tv_sec - st->opened.tv_sec) + 1e-9 * (double)(t1.tv_nsec - st->opened.tv_nsec)), 1e3 * sum);
        for (int k = 0; k < 6; k++) fprintf(stderr, " %s %.1f ms / %ld;", names[k], 1e3 * st->secs[k], st->calls[k]);
        fprintf(stderr, "\n");
    }
    for (std::map<xr::Lowres*, x265cu::Lowres*>::iterator s = st->shadows.begin(); s != st->shadows.end(); ++s)
        st->la.freeLowres(s->second);
    st->shadows.clear();
    st->la.destroy();
    delete st;
}

/* totals over every context closed so far (bench.py: bytes per step, launches per step) */
extern "C" void x265glue_totals(long long* out8)
{
    pthread_mutex_lock(&g_lock);
    memcpy(out8, g_totals, sizeof(g_totals));
    pthread_mutex_unlock(&g_lock);
}

/* ================================================================================================ pre-lookahead */
extern "C" void x265glue_pre_list(xr::Lookahead* la, xr::Frame** frames, int n)
{
    GlueState* st = stateOf(la);
    Timed timed(st, T_PRE);
    std::vector<x265cu::Lowres*> ls((size_t)n);
    std::vector<x265cu::Lookahead::PictureIn> pics((size_t)n);
    for (int i = 0; i < n; i++)
    {
        xr::Frame* f = frames[i];
        /* the per-frame resets of Lowres::init (lowres.cpp:130-153); its pixel work is skipped in this build */
        f->m_lowres.init(f->m_fencPic, f->m_poc);
        ls[i] = shadowOf(st, la, &f->m_lowres);
        xr::PicYuv* pic = f->m_fencPic;
        x265cu::Lookahead::PictureIn in = { pic->m_picOrg[0], pic->m_stride, pic->m_picOrg[1], pic->m_picOrg[2], pic->m_strideC, f->m_poc, f->m_quantOffsets };
        pics[i] = in;
    }
    /* ONE pipelined call for the list: uploads, lowres + variance kernels, float AQ mapping (callback, host), intra */
    /* the lowres planes stay on the device: their only host-side reader was weightPrediction.cpp, whose pixel loops run
     * there too (x265glue_wp_*); the trace harness checksums them, so it gets them */
    const bool planesBack = traceLevel() > 0 || (getenv("X265CU_GLUE_PLANES_BACK") && atoi(getenv("X265CU_GLUE_PLANES_BACK")));
    if (!st->la.preLookaheadBatch(n, &ls[0], &pics[0], planesBack)) die("preLookaheadBatch", st->la.m_error);
    for (int i = 0; i < n; i++)
    {
        xr::Lowres& xl = frames[i]->m_lowres;
        const x265cu::Lowres* sh = ls[i];
        xl.costEst[0][0] = sh->costEst[0][0];
        xl.costEstAq[0][0] = sh->costEstAq[0][0];
        for (int k = 0; k < 3; k++) { xl.wp_ssd[k] = sh->wp_ssd[k]; xl.wp_sum[k] = sh->wp_sum[k]; }
        xl.frameVariance = sh->frameVariance;
        frames[i]->m_lowresInit = true;
    }
    if (traceLevel() && x265ref_hook_pre)
    {
        if (!st->la.sync()) die("x265cu_sync", x265cu_last_error(st->la.m_ctx));     /* the hook checksums the planes */
        for (int i = 0; i < n; i++) x265ref_hook_pre(frames[i]);
    }
}

/* ================================================================================================ estimates */
extern "C" void x265glue_ensure(xr::Lookahead* la, xr::Lowres** frames, int p0, int p1, int b)
{
    xr::Lowres* fenc = frames[b];
    const int d0 = b - p0, d1 = p1 - b;
    if (fenc->costEst[d0][d1] >= 0 && fenc->rowSatds[d0][d1][0] != -1)
        return;                                         /* cached (slicetype.cpp:1982) */
    if (p0 == b) die("estimateFrameCost", "I frame estimates should always be pre-calculated");
    GlueState* st = stateOf(la);
    Timed timed(st, T_SINGLE);
    const int s0 = p0 < b && fenc->lowresMvs[0][d0 - 1][0].x == 0x7FFF;
    const int s1 = p1 > b && fenc->lowresMvs[1][d1 - 1][0].x == 0x7FFF;
    std::vector<x265cu::Lowres*> fr((size_t)(d0 + d1 + 1), (x265cu::Lowres*)NULL);
    fr[0] = knownShadow(st, la, frames[p0]);
    fr[d0 + d1] = knownShadow(st, la, frames[p1]);
    fr[d0] = knownShadow(st, la, fenc);
    x265cu::CostEstimateGroup grp(st->la, &fr[0]);
    if (grp.singleCost(0, d0 + d1, d0, false) < 0) die("singleCost", st->la.m_error);
    mirrorEstimate(fenc, fr[d0], d0, d1);
    traceJob(frames, fr[d0], p0, p1, b, s0, s1, 0, la->m_numCoopSlices > 1 && (p1 > b || s0 || s1));
}

extern "C" int x265glue_finish_batch(xr::Lookahead* la, xr::Lowres** frames, const int* est, int n)
{
    if (traceLevel() && x265ref_hook_batch) x265ref_hook_batch(1, n);
    if (n > 0)
    {
        GlueState* st = stateOf(la);
        Timed timed(st, T_BATCH);
        int lo = est[0], hi = est[2];
        for (int i = 0; i < n; i++) { lo = X265_MIN(lo, est[3 * i]); hi = X265_MAX(hi, est[3 * i + 2]); }
        std::vector<x265cu::Lowres*> fr((size_t)(hi - lo + 1), (x265cu::Lowres*)NULL);
        std::vector<int> state((size_t)n);              /* bit 0/1: searches L0/L1, bit 2: was cached */
        x265cu::CostEstimateGroup* grp = new x265cu::CostEstimateGroup(st->la, &fr[0]);
        for (int i = 0; i < n; i++)
        {
            const int p0 = est[3 * i], b = est[3 * i + 1], p1 = est[3 * i + 2], d0 = b - p0, d1 = p1 - b;
            xr::Lowres* fenc = frames[b];
            if (!fr[p0 - lo]) fr[p0 - lo] = knownShadow(st, la, frames[p0]);
            if (!fr[p1 - lo]) fr[p1 - lo] = knownShadow(st, la, frames[p1]);
            if (!fr[b - lo]) fr[b - lo] = knownShadow(st, la, fenc);
            state[i] = (p0 < b && fenc->lowresMvs[0][d0 - 1][0].x == 0x7FFF ? 1 : 0) | (p1 > b && fenc->lowresMvs[1][d1 - 1][0].x == 0x7FFF ? 2 : 0) |
                       (fenc->costEst[d0][d1] >= 0 && fenc->rowSatds[d0][d1][0] != -1 ? 4 : 0);
            grp->add(p0 - lo, p1 - lo, b - lo);
        }
        /* the whole batch -- every frame pair and B-frame candidate x265 queued -- is ONE x265cu_estimate_batch */
        if (!grp->finishBatch()) die("finishBatch", st->la.m_error);
        delete grp;
        for (int i = 0; i < n; i++)
        {
            if (state[i] & 4) continue;
            const int p0 = est[3 * i], b = est[3 * i + 1], p1 = est[3 * i + 2];
            mirrorEstimate(frames[b], fr[b - lo], b - p0, p1 - b);
            traceJob(frames, fr[b - lo], p0, p1, b, state[i] & 1, (state[i] >> 1) & 1, 1, 0);
        }
    }
    if (traceLevel() && x265ref_hook_batch) x265ref_hook_batch(0, n);
    return 1;
}

/* ================================================================================================ cuTree */
extern "C" void x265glue_ct_zero(xr::Lookahead* la, xr::Lowres* frame)
{
    GlueState* st = stateOf(la);
    Timed timed(st, T_OTHER);
    st->la.cuTreeZero(*shadowOf(st, la, frame));
    if (traceLevel() && x265ref_hook_ctzero) x265ref_hook_ctzero(frame);
}

extern "C" int x265glue_propagate(xr::Lookahead* la, xr::Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced)
{
    GlueState* st = stateOf(la);
    Timed timed(st, T_OTHER);
    const int d0 = b - p0, d1 = p1 - b;
    std::vector<x265cu::Lowres*> fr((size_t)(d0 + d1 + 1), (x265cu::Lowres*)NULL);
    fr[0] = knownShadow(st, la, frames[p0]);
    fr[d0 + d1] = knownShadow(st, la, frames[p1]);
    fr[d0] = knownShadow(st, la, frames[b]);
    /* queued: the steps of a cuTree pass run as one launch when its cuTreeFinish asks for the result */
    if (!st->la.estimateCUPropagate(&fr[0], averageDuration, 0, d0 + d1, d0, referenced)) die("estimateCUPropagate", st->la.m_error);
    if (traceLevel() >= 2)
    {
        /* full trace: the state of the arrays after EVERY step is checksummed, so every step is run and fetched now */
        if (!st->la.propagateCost(*fr[0]) || !st->la.propagateCost(*fr[d0 + d1]) || !st->la.propagateCost(*fr[d0]))
            die("propagateCost", st->la.m_error);
    }
    if (traceLevel() && x265ref_hook_propagate && traceLevel() >= 2) x265ref_hook_propagate(frames, averageDuration, p0, p1, b, referenced);
    return 1;
}

extern "C" void x265glue_ct_fetch(xr::Lookahead* la, xr::Lowres* frame)
{
    GlueState* st = stateOf(la);
    Timed timed(st, T_CTFETCH);
    if (!st->la.propagateCost(*shadowOf(st, la, frame))) die("propagateCost", st->la.m_error);
}

/* Lookahead::cuTreeFinish (slicetype.cpp:1844-1862) through the host layer: the queued pass runs (one launch), the frame's
 * propagateCost comes back, and the log2 mapping to qpCuTreeOffset is x265cu::Lookahead::cuTreeFinish -- the reference's
 * expressions, compiled like the reference, with X265_LOG2 of the (integer) arguments memoised: the same libm call's
 * result, reused (the loop is 0.2 ms per 1080p frame with two log2 calls per CU, 31 times per 60 frames).  Returns 1:
 * the caller returns; 0 (X265CU_GLUE_OWN_FINISH=0): only the fetch was done and x265's own loop runs. */
extern "C" int x265glue_ct_finish(xr::Lookahead* la, xr::Lowres* frame, double averageDuration, int ref0Distance)
{
    static const bool own = !(getenv("X265CU_GLUE_OWN_FINISH") && atoi(getenv("X265CU_GLUE_OWN_FINISH")) == 0);
    if (!own) { x265glue_ct_fetch(la, frame); return 0; }
    GlueState* st = stateOf(la);
    {
        Timed timed(st, T_CTFETCH);
        x265cu::Lowres* sh = shadowOf(st, la, frame);
        if (ref0Distance) sh->weightedCostDelta[ref0Distance - 1] = frame->weightedCostDelta[ref0Distance - 1];
        if (!st->la.cuTreeFinish(sh, averageDuration, ref0Distance)) die("cuTreeFinish", st->la.m_error);
    }
    x265glue_ct_finished(la, frame, averageDuration, ref0Distance);
    return 1;
}

extern "C" void x265glue_ct_finished(xr::Lookahead*, xr::Lowres* frame, double averageDuration, int ref0Distance)
{
    if (traceLevel() && x265ref_hook_ctfinish) x265ref_hook_ctfinish(frame, averageDuration, ref0Distance);
}

/* rc-lookahead 0 (slicetype.cpp:1663-1675, 1732-1737): cuTree swaps the propagateCost POINTERS of two frames.  The device
 * keeps one accumulator per frame slot, so both arrays are brought to the host before the swap and sent back after it. */
extern "C" void x265glue_ct_preswap(xr::Lookahead* la, xr::Lowres* a, xr::Lowres* b)
{
    GlueState* st = stateOf(la);
    if (!st->la.propagateCost(*shadowOf(st, la, a)) || !st->la.propagateCost(*shadowOf(st, la, b))) die("propagateCost", st->la.m_error);
}

extern "C" void x265glue_ct_postswap(xr::Lookahead* la, xr::Lowres* a, xr::Lowres* b)
{
    GlueState* st = stateOf(la);
    xr::Lowres* two[2] = { a, b };
    for (int i = 0; i < 2; i++)
    {
        x265cu::Lowres* sh = shadowOf(st, la, two[i]);      /* re-points the shadow at the swapped array */
        if (x265cu_frame_set_propagate(st->la.m_ctx, sh->slot, sh->propagateCost)) die("x265cu_frame_set_propagate", x265cu_last_error(st->la.m_ctx));
        sh->propagateStale = false;
    }
}

/* ================================================================================================ weightAnalyse */
namespace {
__thread GlueState* t_wpState;      /* the context this thread's weightAnalyse is using (holds its wpLock) */
}

extern "C" void x265glue_wp_done(void)
{
    if (t_wpState) { pthread_mutex_unlock(&t_wpState->wpLock); t_wpState = NULL; }
}

extern "C" void x265glue_wp_prepare(xr::Frame* frame, xr::Frame* refFrame, int plane, const void* mvs)
{
    x265glue_wp_done();
    GlueState* st = NULL;
    x265cu::Lowres *sf = NULL, *sr = NULL;
    pthread_mutex_lock(&g_lock);
    for (std::map<xr::Lookahead*, GlueState*>::iterator it = g_states.begin(); it != g_states.end() && !st; ++it)
    {
        MapGuard guard(&it->second->mapLock);
        std::map<xr::Lowres*, x265cu::Lowres*>::iterator f = it->second->shadows.find(&frame->m_lowres), r = it->second->shadows.find(&refFrame->m_lowres);
        if (f != it->second->shadows.end() && r != it->second->shadows.end()) { st = it->second; sf = f->second; sr = r->second; }
    }
    pthread_mutex_unlock(&g_lock);
    if (!st) die("weightAnalyse", "a frame whose lowres planes are no longer on the device (raise X265CU_FRAME_SLOTS)");
    pthread_mutex_lock(&st->wpLock);
    t_wpState = st;
    if (x265cu_wp_prepare(st->la.m_ctx, sf->slot, sr->slot, plane, mvs, plane == 0 ? frame->m_lowres.intraCost : NULL))
        die("x265cu_wp_prepare", x265cu_last_error(st->la.m_ctx));
}

extern "C" int x265glue_wp_cost(int weighted, int scale, int denom, int offset, unsigned int* cost)
{
    if (!t_wpState) die("weightCost", "no x265glue_wp_prepare before it");
    x265cu_weight_item it = { 0, 0, weighted, scale, denom, offset };
    uint32_t c = 0;
    if (x265cu_wp_cost(t_wpState->la.m_ctx, 1, &it, &c)) die("x265cu_wp_cost", x265cu_last_error(t_wpState->la.m_ctx));
    *cost = c;
    return 1;
}

extern "C" void x265glue_sync(xr::Lookahead* la)
{
    GlueState* st = stateOf(la);
    Timed timed(st, T_SYNC);
    if (!st->la.sync()) die("x265cu_sync", x265cu_last_error(st->la.m_ctx));
}
