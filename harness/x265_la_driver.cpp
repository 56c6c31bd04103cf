/* x265_la_driver.cpp -- lookahead-only driver of x265 1.9 + observation hooks (HARNESS: tests and bench.py).
 *
 * Compiled WITH the reference's headers and linked with the reference's objects into two libraries:
 *   oracle/_ref/libx265ref<depth>.so   the UNMODIFIED reference (slicetype.cpp with observation call-outs only,
 *                                      oracle/make_hooked_slicetype.py): golden traces, `bench.py --impl reference`
 *   oracle/_ref/libx265gpu<depth>.so   the same x265 objects, except that encoder/slicetype.cpp, common/lowres.cpp and
 *                                      common/picyuv.cpp are compiled from temporary copies bound to libx265cu.so
 *                                      (integration/make_gpu_sources.py + integration/x265_glue.cpp):
 *                                      x265's OWN Lookahead (slicetypeDecide, scenecut, slicetypePath, cuTree control flow,
 *                                      vbvLookahead, thread pool) is the host, the cost estimation runs on the GPU.
 * Both export the same entry points, so the same driver, the same clip and the same pool size time both arms
 * (`bench.py`: e2e = libx265gpu, `--impl reference` = libx265ref) and the same trace is written by both (parity).
 *
 *   x265la_open   params, thread pool, all input Frames (synthetic clip generated and copied into PicYuv)
 *   x265la_run    a fresh Lookahead over those Frames: addPicture -> getDecidedPicture for every frame, flush;
 *                 returns the wall seconds between the first addPicture and the last decided picture
 *                 (bring-up order of x265_encoder_open / Encoder::create, api.cpp:94-99, encoder.cpp:130-209,316)
 *   x265la_close
 *   x265ref_run_lookahead   open + run + close (kept for the golden-trace generator)
 * The hooks write a TRACE of every estimate the lookahead ran with CRC32s of all outputs and optionally a binary
 * DUMP of the arrays; they only observe.
 */
#include "common.h"
#include "primitives.h"
#include "param.h"
#include "frame.h"
#include "picyuv.h"
#include "lowres.h"
#include "slicetype.h"
#include "threadpool.h"
#include "encoder.h"
#include "bitcost.h"
#include "motion.h"
#include "x265.h"

#include "../oracle/ref_hooks.h"
#include "../oracle/synth.h"

#include <stdio.h>
#include <string.h>
#include <time.h>
#include <map>
#include <vector>
#include <pthread.h>

using namespace X265_NS;

namespace {

/* ---------------------------------------------------------------- CRC32 (zlib polynomial) */
uint32_t g_crcTab[256];
void crcInit()
{
    if (g_crcTab[1]) return;
    for (uint32_t i = 0; i < 256; i++)
    {
        uint32_t c = i;
        for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
        g_crcTab[i] = c;
    }
}
uint32_t crc32(const void* p, size_t n, uint32_t crc = 0)
{
    const uint8_t* b = (const uint8_t*)p;
    crc = ~crc;
    while (n--) crc = g_crcTab[(crc ^ *b++) & 255] ^ (crc >> 8);
    return ~crc;
}

/* ---------------------------------------------------------------- trace state */
struct TraceState
{
    FILE* trace;
    FILE* dump;
    pthread_mutex_t lock;
    int wCU, hCU, nCU, bframes;
    long nPre, nJob, nSearch[2], nBatch, nPropagate;
} g_ts;
std::map<std::pair<int, int>, std::vector<int> > g_weights; /* (fencPoc, refPoc) -> scale, denom, offset */

void dumpArray(const char tag[4], int a, int b, int c, int id, const void* data, size_t bytes)
{
    if (!g_ts.dump) return;
    int32_t hdr[6] = { 0, a, b, c, id, (int32_t)bytes };
    memcpy(&hdr[0], tag, 4);
    fwrite(hdr, sizeof(hdr), 1, g_ts.dump);
    fwrite(data, 1, bytes, g_ts.dump);
}

struct ExposeBitCost : public BitCost
{
    const uint16_t* table() const { return m_cost; }
};

} // namespace

/* ================================================================ hooks */
extern "C" void x265ref_hook_pre(Frame* frame)
{
    Lowres& l = frame->m_lowres;
    pthread_mutex_lock(&g_ts.lock);
    if (g_ts.trace)
    {
        size_t planeBytes = (size_t)(l.buffer[1] - l.buffer[0]) * sizeof(pixel);
        int n = g_ts.nCU;
        uint32_t cPlanes = crc32(l.buffer[0], 4 * planeBytes);
        uint32_t cIC = crc32(l.intraCost, n * sizeof(int32_t));
        uint32_t cIM = crc32(l.intraMode, n);
        uint32_t cLC = crc32(l.lowresCosts[0][0], n * sizeof(uint16_t));
        uint32_t cRS = crc32(l.rowSatds[0][0], g_ts.hCU * sizeof(int32_t));
        uint32_t cIQ = l.invQscaleFactor ? crc32(l.invQscaleFactor, n * sizeof(int)) : 0;
        fprintf(g_ts.trace, "P %d %08x %08x %08x %08x %08x %lld %lld %08x %llu %llu\n", l.frameNum, cPlanes, cIC, cIM, cLC, cRS,
                (long long)l.costEst[0][0], (long long)l.costEstAq[0][0], cIQ,
                (unsigned long long)l.wp_ssd[0], (unsigned long long)l.wp_sum[0]);
        if (l.frameNum == 0)
            dumpArray("PLAN", l.frameNum, 0, 0, 0, l.buffer[0], 4 * planeBytes);
        dumpArray("ICST", l.frameNum, 0, 0, 0, l.intraCost, n * sizeof(int32_t));
        dumpArray("IMOD", l.frameNum, 0, 0, 0, l.intraMode, n);
        dumpArray("LCST", l.frameNum, l.frameNum, l.frameNum, 0, l.lowresCosts[0][0], n * sizeof(uint16_t));
        dumpArray("RSAT", l.frameNum, l.frameNum, l.frameNum, 0, l.rowSatds[0][0], g_ts.hCU * sizeof(int32_t));
        if (l.invQscaleFactor)
            dumpArray("INVQ", l.frameNum, 0, 0, 0, l.invQscaleFactor, n * sizeof(int));
    }
    g_ts.nPre++;
    pthread_mutex_unlock(&g_ts.lock);
}

extern "C" void x265ref_hook_batch(int begin, int njobs)
{
    pthread_mutex_lock(&g_ts.lock);
    if (g_ts.trace)
        fprintf(g_ts.trace, begin ? "B %d\n" : "E %d\n", njobs);
    if (begin) g_ts.nBatch++;
    pthread_mutex_unlock(&g_ts.lock);
}

extern "C" void x265ref_hook_weight(int fencPoc, int refPoc, int scale, int denom, int offset)
{
    pthread_mutex_lock(&g_ts.lock);
    std::vector<int> w(3);
    w[0] = scale; w[1] = denom; w[2] = offset;
    g_weights[std::make_pair(fencPoc, refPoc)] = w;
    pthread_mutex_unlock(&g_ts.lock);
}

extern "C" void x265ref_hook_job(Lowres** frames, int p0, int p1, int b, int search0, int search1, int batchMode, int sliced)
{
    Lowres* fenc = frames[b];
    int d0 = b - p0, d1 = p1 - b;
    int n = g_ts.nCU;
    pthread_mutex_lock(&g_ts.lock);
    g_ts.nJob++;
    g_ts.nSearch[0] += search0;
    g_ts.nSearch[1] += search1;
    if (g_ts.trace)
    {
        int pocB = fenc->frameNum, poc0 = frames[p0]->frameNum, poc1 = frames[p1]->frameNum;
        int wflag = 0, ws = 0, wd = 0, wo = 0;
        if (search0 && fenc->weightedRef[d0].isWeighted)
        {
            std::map<std::pair<int, int>, std::vector<int> >::iterator it = g_weights.find(std::make_pair(pocB, poc0));
            if (it != g_weights.end())
            {
                wflag = 1; ws = it->second[0]; wd = it->second[1]; wo = it->second[2];
            }
            else
                wflag = -1; /* should not happen */
        }
        uint32_t cMv[2] = { 0, 0 }, cMc[2] = { 0, 0 };
        if (d0 > 0)
        {
            cMv[0] = crc32(fenc->lowresMvs[0][d0 - 1], n * sizeof(MV));
            cMc[0] = crc32(fenc->lowresMvCosts[0][d0 - 1], n * sizeof(int32_t));
        }
        if (d1 > 0)
        {
            cMv[1] = crc32(fenc->lowresMvs[1][d1 - 1], n * sizeof(MV));
            cMc[1] = crc32(fenc->lowresMvCosts[1][d1 - 1], n * sizeof(int32_t));
        }
        uint32_t cLC = crc32(fenc->lowresCosts[d0][d1], n * sizeof(uint16_t));
        uint32_t cRS = crc32(fenc->rowSatds[d0][d1], g_ts.hCU * sizeof(int32_t));
        /* weightedCostDelta is a double written by weightsAnalyse (slicetype.cpp:472) */
        double wdelta = fenc->weightedCostDelta[d0];
        fprintf(g_ts.trace, "J %d %d %d %d %d %d %d %d %d %d %d %lld %lld %d %08x %08x %08x %08x %08x %08x %.17g\n",
                poc0, pocB, poc1, search0, search1, batchMode, sliced, wflag, ws, wd, wo,
                (long long)fenc->costEst[d0][d1], (long long)fenc->costEstAq[d0][d1], fenc->intraMbs[d0],
                cMv[0], cMv[1], cMc[0], cMc[1], cLC, cRS, wdelta);
        if (g_ts.dump)
        {
            if (d0 > 0 && search0)
            {
                dumpArray("MVS0", poc0, pocB, poc1, d0, fenc->lowresMvs[0][d0 - 1], n * sizeof(MV));
                dumpArray("MVC0", poc0, pocB, poc1, d0, fenc->lowresMvCosts[0][d0 - 1], n * sizeof(int32_t));
            }
            if (d1 > 0 && search1)
            {
                dumpArray("MVS1", poc0, pocB, poc1, d1, fenc->lowresMvs[1][d1 - 1], n * sizeof(MV));
                dumpArray("MVC1", poc0, pocB, poc1, d1, fenc->lowresMvCosts[1][d1 - 1], n * sizeof(int32_t));
            }
            dumpArray("LCST", poc0, pocB, poc1, 0, fenc->lowresCosts[d0][d1], n * sizeof(uint16_t));
            dumpArray("RSAT", poc0, pocB, poc1, 0, fenc->rowSatds[d0][d1], g_ts.hCU * sizeof(int32_t));
        }
    }
    pthread_mutex_unlock(&g_ts.lock);
}

/* ---- cuTree: which arrays Lookahead::cuTree zeroed, every propagate step with the state of the
 * reference frames' propagateCost after it, every cuTreeFinish with the qpCuTreeOffset it wrote */
extern "C" void x265ref_hook_ctzero(Lowres* frame)
{
    pthread_mutex_lock(&g_ts.lock);
    if (g_ts.trace) fprintf(g_ts.trace, "M %d\n", frame->frameNum);
    pthread_mutex_unlock(&g_ts.lock);
}

extern "C" void x265ref_hook_propagate(Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced)
{
    pthread_mutex_lock(&g_ts.lock);
    if (g_ts.trace)
    {
        int n = g_ts.nCU;
        uint32_t c0 = crc32(frames[p0]->propagateCost, n * sizeof(uint16_t));
        uint32_t c1 = p1 != b ? crc32(frames[p1]->propagateCost, n * sizeof(uint16_t)) : 0;
        /* a non-referenced frame only has its first row defined (zeroed by the step itself) */
        uint32_t cb = referenced ? crc32(frames[b]->propagateCost, n * sizeof(uint16_t)) : crc32(frames[b]->propagateCost, g_ts.wCU * sizeof(uint16_t));
        fprintf(g_ts.trace, "X %d %d %d %d %.17g %08x %08x %08x\n", frames[p0]->frameNum, frames[b]->frameNum, frames[p1]->frameNum, referenced,
                averageDuration, c0, c1, cb);
        dumpArray("PRP0", frames[p0]->frameNum, frames[b]->frameNum, frames[p1]->frameNum, referenced, frames[p0]->propagateCost, n * sizeof(uint16_t));
        if (p1 != b)
            dumpArray("PRP1", frames[p0]->frameNum, frames[b]->frameNum, frames[p1]->frameNum, referenced, frames[p1]->propagateCost, n * sizeof(uint16_t));
    }
    g_ts.nPropagate++;
    pthread_mutex_unlock(&g_ts.lock);
}

extern "C" void x265ref_hook_ctfinish(Lowres* frame, double averageDuration, int ref0Distance)
{
    pthread_mutex_lock(&g_ts.lock);
    if (g_ts.trace)
    {
        int n = g_ts.nCU;
        fprintf(g_ts.trace, "F %d %d %.17g %08x %08x\n", frame->frameNum, ref0Distance, averageDuration,
                crc32(frame->qpCuTreeOffset, n * sizeof(double)), crc32(frame->propagateCost, n * sizeof(uint16_t)));
        dumpArray("QPCT", frame->frameNum, ref0Distance, 0, 0, frame->qpCuTreeOffset, n * sizeof(double));
    }
    pthread_mutex_unlock(&g_ts.lock);
}


/* ================================================================ the driver */
namespace {

struct Driver
{
    x265_param* p;
    Encoder* cfg;
    int width, height, nframes, poolThreads;
    uint32_t seed;
    std::vector<Frame*> frames;
};

void driverSetup()
{
    crcInit();
    static bool done = false;
    if (done) return;
    x265_param* sp = x265_param_alloc();
    x265_param_default(sp);
    sp->cpuid = 0;
    sp->logLevel = X265_LOG_NONE;
    x265_setup_primitives(sp);
    MotionEstimate::initScales();
    done = true;
}

void mvcostTable(uint16_t* out)
{
    ExposeBitCost bc;
    bc.setQP(X265_LOOKAHEAD_QP);
    memcpy(out, bc.table() - 2 * 32768, (4 * 32768 + 1) * sizeof(uint16_t));
}

} // namespace

extern "C" {

/* 0: no trace, 1: everything but the per-step propagateCost CRCs (X lines; a GPU host that queues its cuTree steps
 * would have to fetch after every step to show them), 2: full */
int g_x265la_traceLevel = 0;
int x265la_trace_level(void) { return g_x265la_traceLevel; }

/* opts: name/value pairs for x265_param_parse (value may be NULL); "preset"/"tune" are consumed first.
 * poolThreads: size of the single worker pool (results depend on it). */
void* x265la_open(int width, int height, int nframes, uint32_t seed, const char** optNames, const char** optValues, int nopts, int poolThreads, int* err)
{
    int dummy;
    if (!err) err = &dummy;
    *err = 0;
    driverSetup();
    x265_param* p = x265_param_alloc();
    const char* preset = "medium";
    const char* tune = NULL;
    for (int i = 0; i < nopts; i++)
    {
        if (!strcmp(optNames[i], "preset")) preset = optValues[i];
        if (!strcmp(optNames[i], "tune")) tune = optValues[i];
    }
    if (x265_param_default_preset(p, preset, tune) < 0) { *err = -1; return NULL; }
    p->sourceWidth = width;
    p->sourceHeight = height;
    p->fpsNum = 30; p->fpsDenom = 1;
    p->internalCsp = X265_CSP_I420;
    p->logLevel = X265_LOG_NONE;
    p->cpuid = 0;
    p->totalFrames = nframes;
    for (int i = 0; i < nopts; i++)
    {
        if (!strcmp(optNames[i], "preset") || !strcmp(optNames[i], "tune")) continue;
        if (x265_param_parse(p, optNames[i], optValues[i]) < 0) { fprintf(stderr, "x265la: bad option %s\n", optNames[i]); *err = -2; return NULL; }
    }
    x265_setup_primitives(p);
    if (x265_check_params(p)) { *err = -3; return NULL; }
    if (x265_set_globals(p)) { *err = -4; return NULL; }
    Encoder* cfg = new Encoder;
    cfg->configure(p);     /* normalises lookahead/bframes/slices parameters exactly as the encoder does */
    if (p->sourceWidth != width || p->sourceHeight != height) { *err = -5; return NULL; } /* only unpadded sizes supported */

    Driver* d = new Driver;
    d->p = p; d->cfg = cfg;
    d->width = width; d->height = height; d->nframes = nframes; d->poolThreads = poolThreads; d->seed = seed;
    if (poolThreads <= 0)
        p->lookaheadSlices = 0;

    /* generate all input frames up front */
    d->frames.resize(nframes);
    std::vector<pixel> y((size_t)width * height), u((size_t)width * height / 4), v((size_t)width * height / 4);
    x265_picture pic;
    x265_picture_init(p, &pic);
    pic.bitDepth = X265_DEPTH;
    pic.planes[0] = &y[0]; pic.planes[1] = &u[0]; pic.planes[2] = &v[0];
    pic.stride[0] = width * (int)sizeof(pixel);
    pic.stride[1] = pic.stride[2] = (width / 2) * (int)sizeof(pixel);
    for (int t = 0; t < nframes; t++)
    {
        synth_frame(width, height, t, nframes, seed, X265_DEPTH, &y[0], width, &u[0], &v[0], width / 2);
        Frame* f = new Frame;
        if (!f->create(p, NULL)) { *err = -8; return NULL; }
        f->m_fencPic->copyFromPicture(pic, *p, 0, 0);
        f->m_poc = t;
        d->frames[t] = f;
    }
    return d;
}

/* one pass of a fresh Lookahead over the clip.  sliceTypesOut[nframes]: decided X265_TYPE_* per POC.  statsOut[8]: nPre,
 * nJobs, nSearchL0, nSearchL1, nBatches, decided, nPropagate, 0.  Returns wall seconds (Lookahead construction / create()
 * and the input pictures are outside the clock, for either arm), < 0 on error. */
double x265la_run(void* h, const char* tracePath, const char* dumpPath, int traceLevel, int* sliceTypesOut, long* statsOut)
{
    Driver* d = (Driver*)h;
    x265_param* p = d->p;
    const int nframes = d->nframes;
    /* a pool per run: its workers keep pointers to the job providers they served */
    ThreadPool* pool = NULL;
    if (d->poolThreads > 0)
    {
        pool = new ThreadPool[1];
        if (!pool[0].create(d->poolThreads, 1, 0)) return -6;
    }

    const int lookaheadSlices = p->lookaheadSlices;    /* the constructor replaces it by the slice count it settled on */
    Lookahead* la = new Lookahead(p, pool);
    p->lookaheadSlices = lookaheadSlices;
    if (pool)
    {
        la->m_jpId = pool[0].m_numProviders++;
        pool[0].m_jpTable[la->m_jpId] = la;
        pool[0].start();
    }
    if (!la->create()) return -7;

    memset(&g_ts, 0, sizeof(g_ts));
    g_weights.clear();
    pthread_mutex_init(&g_ts.lock, NULL);
    g_ts.wCU = la->m_8x8Width; g_ts.hCU = la->m_8x8Height; g_ts.nCU = g_ts.wCU * g_ts.hCU; g_ts.bframes = p->bframes;
    g_ts.trace = tracePath ? fopen(tracePath, "w") : NULL;
    g_ts.dump = dumpPath ? fopen(dumpPath, "wb") : NULL;
    g_x265la_traceLevel = g_ts.trace ? (traceLevel > 0 ? traceLevel : 2) : 0;
    if (g_ts.trace)
    {
        fprintf(g_ts.trace, "# x265la-trace v1 (reference x265 1.9, %d-bit, C primitives)\n", X265_DEPTH);
        fprintf(g_ts.trace, "C %d %d %d %d %u %d %d %d %d %d %d %d %d %d %d\n", d->width, d->height, X265_DEPTH, nframes, d->seed,
                p->bframes, p->lookaheadDepth, p->bFrameAdaptive, p->bEnableWeightedPred, p->rc.aqMode, p->rc.cuTree,
                la->m_numCoopSlices, la->m_numRowsPerSlice, p->bFrameBias, d->poolThreads);
        fprintf(g_ts.trace, "Q %.17g %d %d\n", p->rc.aqStrength, p->scenecutThreshold, p->keyframeMax);
        {
            std::vector<uint16_t> lut(4 * 32768 + 1);
            mvcostTable(&lut[0]);
            fprintf(g_ts.trace, "L %08x %d\n", crc32(&lut[0], lut.size() * sizeof(uint16_t)), (int)x265_lambda_tab[X265_LOOKAHEAD_QP]);
        }
        /* what cuTree reads beyond the above: frame rate, cuTree strength (5 * (1 - qcomp)), weighted bipred, VBV */
        fprintf(g_ts.trace, "T %u %u %.17g %d %d\n", p->fpsNum, p->fpsDenom, p->rc.qCompress, p->bEnableWeightedBiPred, p->rc.vbvBufferSize);
    }

    /* the per-picture state Encoder::encode sets before Lookahead::addPicture (encoder.cpp:582-584) */
    for (int t = 0; t < nframes; t++)
    {
        Frame* f = d->frames[t];
        f->m_lowres.bScenecut = false;
        f->m_lowres.satdCost = (int64_t)-1;
        f->m_lowresInit = false;
        f->m_next = f->m_prev = NULL;
    }

    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    int decided = 0;
    for (int t = 0; t < nframes; t++)
    {
        la->addPicture(*d->frames[t], X265_TYPE_AUTO);
        Frame* out = la->getDecidedPicture();
        if (out)
        {
            if (sliceTypesOut) sliceTypesOut[out->m_poc] = out->m_lowres.sliceType;
            if (g_ts.trace) { pthread_mutex_lock(&g_ts.lock); fprintf(g_ts.trace, "D %d %d\n", out->m_poc, out->m_lowres.sliceType); pthread_mutex_unlock(&g_ts.lock); }
            decided++;
        }
    }
    la->flush();
    while (decided < nframes)
    {
        Frame* out = la->getDecidedPicture();
        if (!out) break;
        if (sliceTypesOut) sliceTypesOut[out->m_poc] = out->m_lowres.sliceType;
        if (g_ts.trace) { pthread_mutex_lock(&g_ts.lock); fprintf(g_ts.trace, "D %d %d\n", out->m_poc, out->m_lowres.sliceType); pthread_mutex_unlock(&g_ts.lock); }
        decided++;
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    double secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);

    if (statsOut)
    {
        statsOut[0] = g_ts.nPre; statsOut[1] = g_ts.nJob; statsOut[2] = g_ts.nSearch[0]; statsOut[3] = g_ts.nSearch[1];
        statsOut[4] = g_ts.nBatch; statsOut[5] = decided; statsOut[6] = g_ts.nPropagate; statsOut[7] = 0;
    }
    if (g_ts.trace) { fprintf(g_ts.trace, "Z %d\n", decided); fclose(g_ts.trace); g_ts.trace = NULL; }
    if (g_ts.dump) { fclose(g_ts.dump); g_ts.dump = NULL; }
    g_x265la_traceLevel = 0;

    la->stopJobs();
    if (pool) { pool->stopWorkers(); delete[] pool; }
    /* every picture has been handed out: both queues are empty, no slicetypeDecide is running */
    la->destroy();
    delete la;
    return decided == nframes ? secs : -9;
}

void x265la_close(void* h)
{
    Driver* d = (Driver*)h;
    if (!d) return;
    for (size_t t = 0; t < d->frames.size(); t++)
        if (d->frames[t]) { d->frames[t]->destroy(); delete d->frames[t]; }
    delete d->cfg;
    x265_param_free(d->p);
    delete d;
}

double x265ref_run_lookahead(int width, int height, int nframes, uint32_t seed,
                             const char** optNames, const char** optValues, int nopts, int poolThreads,
                             const char* tracePath, const char* dumpPath, int* sliceTypesOut, long* statsOut)
{
    int err = 0;
    void* h = x265la_open(width, height, nframes, seed, optNames, optValues, nopts, poolThreads, &err);
    if (!h) return err;
    double secs = x265la_run(h, tracePath, dumpPath, 2, sliceTypesOut, statsOut);
    x265la_close(h);
    return secs;
}

uint32_t x265ref_crc32(const void* p, size_t n) { crcInit(); return crc32(p, n); }

} // extern "C"
