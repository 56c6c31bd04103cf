/* lookahead_cu.cpp -- see lookahead_cu.h.  Host layer above the C ABI; float decisions only. */
#include "lookahead_cu.h"
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include <math.h>
#include <sched.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace x265cu {

/* A small persistent pool for the float AQ mapping (one per process, created on first use): the reference spreads the frames
 * of a pre-lookahead list over its workers (slicetype.cpp:837-855); here the GPU delivers the energies of a few frames at a
 * time, so the frames of such a run AND the block rows inside a frame (independent in aq-mode 1) are spread over the cores
 * this process may use.  run(n, f) calls f(0..n-1), the caller takes part; a second caller at the same time (another stream
 * of the process) simply runs its items itself. */
class HostPool
{
public:
    static HostPool& get() { static HostPool p; return p; }
    int threads() const { return (int)m_threads.size() + 1; }
    void run(int n, const std::function<void(int)>& f)
    {
        if (n <= 1 || m_threads.empty() || !m_busy.try_lock()) { for (int i = 0; i < n; i++) f(i); return; }
        {
            std::lock_guard<std::mutex> lk(m_mtx);
            m_fn = &f; m_total = n; m_next = 0; m_pending = n; m_gen++;
        }
        m_work.notify_all();
        take();
        {
            std::unique_lock<std::mutex> lk(m_mtx);
            m_done.wait(lk, [this]() { return m_pending == 0; });
            m_fn = NULL;
        }
        m_busy.unlock();
    }

private:
    HostPool() : m_fn(NULL), m_total(0), m_next(0), m_pending(0), m_gen(0), m_stop(false)
    {
        unsigned n = 1;
        if (const char* e = getenv("X265CU_HOST_THREADS")) n = (unsigned)atoi(e);
        else
        {
            cpu_set_t set;
            n = std::thread::hardware_concurrency();
            if (sched_getaffinity(0, sizeof(set), &set) == 0 && CPU_COUNT(&set) > 0) n = (unsigned)CPU_COUNT(&set);
            if (n > 16) n = 16;
        }
        for (unsigned t = 1; t < n; t++) m_threads.push_back(std::thread([this]() { loop(); }));
    }
    ~HostPool()
    {
        { std::lock_guard<std::mutex> lk(m_mtx); m_stop = true; }
        m_work.notify_all();
        for (size_t t = 0; t < m_threads.size(); t++) m_threads[t].join();
    }
    void take()
    {
        for (;;)
        {
            int i;
            const std::function<void(int)>* f;
            {
                std::lock_guard<std::mutex> lk(m_mtx);
                if (!m_fn || m_next >= m_total) return;
                i = m_next++; f = m_fn;
            }
            (*f)(i);
            bool last;
            { std::lock_guard<std::mutex> lk(m_mtx); last = --m_pending == 0; }
            if (last) m_done.notify_all();
        }
    }
    void loop()
    {
        uint64_t seen = 0;
        for (;;)
        {
            {
                std::unique_lock<std::mutex> lk(m_mtx);
                m_work.wait(lk, [this, seen]() { return m_stop || (m_gen != seen && m_fn && m_next < m_total); });
                if (m_stop) return;
                seen = m_gen;
            }
            take();
        }
    }
    std::vector<std::thread> m_threads;
    std::mutex m_mtx, m_busy;
    std::condition_variable m_work, m_done;
    const std::function<void(int)>* m_fn;
    int m_total, m_next, m_pending;
    uint64_t m_gen;
    bool m_stop;
};

namespace {
inline int imin(int a, int b) { return a < b ? a : b; }
inline int imax(int a, int b) { return a > b ? a : b; }
inline int iclip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
inline size_t alignUp(size_t v, size_t a) { return (v + a - 1) / a * a; }

/* x265_exp2fix8, common/common.cpp:94-101; LUT entries are round(256 * (2^(i/64) - 1)) */
int exp2fix8(double x)
{
    struct Lut
    {
        uint8_t v[64];
        Lut() { for (int i = 0; i < 64; i++) v[i] = (uint8_t)floor(256.0 * (pow(2.0, i / 64.0) - 1.0) + 0.5); }
    };
    static const Lut table;         /* (thread-safe initialisation: the mapping runs on several threads) */
    const uint8_t* lut = table.v;
    int i = (int)(x * (-64.f / 6.f) + 512.5f);
    if (i < 0) return 0;
    if (i > 1023) return 0xffff;
    return (lut[i & 63] + 256) << (i >> 6) >> 8;
}

uint32_t crc32(const void* p, size_t n)
{
    static uint32_t tab[256];
    if (!tab[1])
        for (uint32_t i = 0; i < 256; i++)
        {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            tab[i] = c;
        }
    const uint8_t* b = (const uint8_t*)p;
    uint32_t crc = 0xFFFFFFFFu;
    while (n--) crc = tab[(crc ^ *b++) & 255] ^ (crc >> 8);
    return ~crc;
}
} // namespace

/* BitCost::CalculateLogs + BitCost::setQP(X265_LOOKAHEAD_QP), encoder/bitcost.cpp:30-59,73-90.
 * lambda = x265_lambda_tab[12 + 6 * (depth - 8)] = 2^(depth - 8) exactly (constants.cpp:74-125). */
void Lookahead::mvcostTable(int bitDepth, uint16_t* out, int* lambdaInt)
{
    const int M = 2 * 32768;
    double lambda = (double)(1 << (2 * (bitDepth - 8)));
    if (lambdaInt) *lambdaInt = (int)lambda;
    float log2_2 = 2.0f / logf(2.0f);
    for (int i = 0; i <= M; i++)
    {
        float bits = i ? logf((float)(i + 1)) * log2_2 + 1.718f : 0.718f;
        double v = bits * lambda + 0.5f;
        if (v > 32767.0) v = 32767.0;
        out[M + i] = out[M - i] = (uint16_t)v;
    }
}

Lookahead::Lookahead() : m_ctx(NULL), m_mvcost(NULL), m_lambda(1), m_resident(false), m_lookAhead(true), m_versionCounter(0), m_newestReady(0), m_episodeNewest(0), m_trellisAhead(true), m_trellisAtBatch(false), m_trellisBref(false), m_batchFirst(-1), m_batchLast(-1)
{
    m_error[0] = 0;
    memset(&m_param, 0, sizeof(m_param));
    memset(m_specStats, 0, sizeof(m_specStats));
    if (const char* e = getenv("X265CU_LOOKAHEAD_CACHE")) m_lookAhead = atoi(e) != 0;
    if (const char* e = getenv("X265CU_TRELLIS_AHEAD")) { m_trellisAhead = (atoi(e) & 1) != 0; m_trellisAtBatch = (atoi(e) & 2) != 0; m_trellisBref = (atoi(e) & 4) != 0; }
}
Lookahead::~Lookahead() { destroy(); }

/* Lookahead::Lookahead + Lookahead::create, encoder/slicetype.cpp:490-591 */
bool Lookahead::create(const Param& p)
{
    m_param = p;
    if (!m_param.maxCUSize) m_param.maxCUSize = 64;
    if (m_param.fpsNum <= 0 || m_param.fpsDenom <= 0) { m_param.fpsNum = 30; m_param.fpsDenom = 1; }
    if (m_param.qCompress <= 0) m_param.qCompress = 0.6;
    m_cuTreeStrength = 5.0 * (1.0 - m_param.qCompress);      /* slicetype.cpp:511 */
    m_ctStats[0] = m_ctStats[1] = m_ctStats[2] = 0;
    m_ctOps.clear();
    m_8x8Height = ((p.sourceHeight / 2) + 7) >> 3;
    m_8x8Width = ((p.sourceWidth / 2) + 7) >> 3;
    m_cuCount = m_8x8Width * m_8x8Height;
    m_8x8Blocks = m_8x8Width > 2 && m_8x8Height > 2 ? (m_cuCount + 4 - 2 * (m_8x8Width + m_8x8Height)) : m_cuCount;
    m_bAdaptiveQuant = p.aqMode || p.bEnableWeightedPred || p.bEnableWeightedBiPred;   /* slicetype.cpp:517 */
    int slices = p.lookaheadSlices;
    if (slices && !p.poolWorkers) slices = 0;
    if (slices && p.sourceHeight < 720) slices = 0;
    if (slices > 1)
    {
        m_numRowsPerSlice = m_8x8Height / slices;
        m_numRowsPerSlice = imax(m_numRowsPerSlice, 10);
        m_numRowsPerSlice = imin(m_numRowsPerSlice, m_8x8Height);
        m_numCoopSlices = m_8x8Height / m_numRowsPerSlice;
    }
    else
    {
        m_numRowsPerSlice = m_8x8Height;
        m_numCoopSlices = 1;
    }
    if (p.forceCoopSlices > 0 && p.forceRowsPerSlice > 0)
    {
        m_numCoopSlices = p.forceCoopSlices;
        m_numRowsPerSlice = p.forceRowsPerSlice;
    }
    m_mvcost = (uint16_t*)malloc((4 * 32768 + 1) * sizeof(uint16_t));
    if (!m_mvcost) return false;
    mvcostTable(p.bitDepth, m_mvcost, &m_lambda);

    x265cu_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.srcWidth = p.sourceWidth; cfg.srcHeight = p.sourceHeight; cfg.bitDepth = p.bitDepth;
    cfg.marginX = m_param.maxCUSize + 32; cfg.marginY = m_param.maxCUSize + 16;   /* picyuv.cpp:62-63 */
    cfg.bframes = p.bframes;
    cfg.numFrameSlots = p.frameSlots > 0 ? p.frameSlots : p.lookaheadDepth + p.bframes + 8;
    cfg.numCoopSlices = m_numCoopSlices; cfg.numRowsPerSlice = m_numRowsPerSlice;
    cfg.bFrameBias = p.bFrameBias;
    cfg.lookaheadLambda = m_lambda;
    cfg.mvcost = m_mvcost + 2 * 32768;
    cfg.device = p.device;
    cfg.stream = p.stream;
    cfg.searchWarps = p.searchWarps;
    int r = x265cu_open(&cfg, &m_ctx);
    if (r != X265CU_OK)
    {
        snprintf(m_error, sizeof(m_error), "x265cu_open failed (%d): %s", r, x265cu_last_error(NULL));
        m_ctx = NULL;
        return false;
    }
    x265cu_get_geometry(m_ctx, &m_geom);
    m_freeSlots.clear();
    for (int i = cfg.numFrameSlots - 1; i >= 0; i--) m_freeSlots.push_back(i);
    return true;
}

void Lookahead::destroy()
{
    if (m_ctx && getenv("X265CU_LOOKAHEAD_STATS"))
        fprintf(stderr, "look-ahead estimate cache: %lld launched ahead, %lld handed out, %lld requests computed on demand, %lld requests\n",
                (long long)m_specStats[0], (long long)m_specStats[1], (long long)m_specStats[2], (long long)m_specStats[3]);
    if (m_ctx) { x265cu_close(m_ctx); m_ctx = NULL; }
    free(m_mvcost); m_mvcost = NULL;
}

/* Lowres::create, common/lowres.cpp:30-95: every output array, carved out of one pinned arena */
Lowres* Lookahead::allocLowres()
{
    if (m_freeSlots.empty()) { snprintf(m_error, sizeof(m_error), "allocLowres: no free frame slot"); return NULL; }
    Lowres* l = (Lowres*)calloc(1, sizeof(Lowres));
    const int bf = m_param.bframes, n = m_cuCount, rows = m_8x8Height;
    const size_t pb = m_geom.pixelBytes;
    size_t bytes = alignUp((size_t)4 * m_geom.planeSize * pb, 64);
    bytes += alignUp((size_t)n * 4, 64) + alignUp((size_t)n, 64) + alignUp((size_t)n * 2, 64);   /* intraCost, intraMode, propagateCost */
    bytes += 2 * alignUp((size_t)n * 8, 64) + 2 * alignUp((size_t)n * 4, 64);             /* qpAq, qpCuTree, invQ, blockVariance */
    bytes += (size_t)(bf + 2) * (bf + 2) * (alignUp((size_t)rows * 4, 64) + alignUp((size_t)n * 2, 64));
    bytes += (size_t)2 * (bf + 1) * 2 * alignUp((size_t)n * 4, 64);
    l->arenaBytes = bytes;
    if (posix_memalign((void**)&l->arena, 4096, bytes)) { free(l); return NULL; }
    memset(l->arena, 0, bytes);
    x265cu_host_register(l->arena, bytes);   /* pinned: async copies, no staging; failure only costs speed */
    uint8_t* p = l->arena;
    for (int i = 0; i < 4; i++)
    {
        l->buffer[i] = p + (size_t)i * m_geom.planeSize * pb;
        l->lowresPlane[i] = (uint8_t*)l->buffer[i] + (size_t)m_geom.padOffset * pb;
    }
    p += alignUp((size_t)4 * m_geom.planeSize * pb, 64);
    l->intraCost = (int32_t*)p; p += alignUp((size_t)n * 4, 64);
    l->intraMode = p; p += alignUp((size_t)n, 64);
    l->propagateCost = (uint16_t*)p; p += alignUp((size_t)n * 2, 64);
    l->qpAqOffset = (double*)p; p += alignUp((size_t)n * 8, 64);
    l->qpCuTreeOffset = (double*)p; p += alignUp((size_t)n * 8, 64);
    l->invQscaleFactor = (int*)p; p += alignUp((size_t)n * 4, 64);
    l->blockVariance = (uint32_t*)p; p += alignUp((size_t)n * 4, 64);
    if (!m_param.aqMode)
        l->qpAqOffset = l->qpCuTreeOffset = NULL, l->invQscaleFactor = NULL, l->blockVariance = NULL;   /* lowres.cpp:53-59 */
    for (int i = 0; i < bf + 2; i++)
        for (int j = 0; j < bf + 2; j++)
        {
            l->rowSatds[i][j] = (int32_t*)p; p += alignUp((size_t)rows * 4, 64);
            l->lowresCosts[i][j] = (uint16_t*)p; p += alignUp((size_t)n * 2, 64);
        }
    for (int i = 0; i < bf + 1; i++)
        for (int k = 0; k < 2; k++)
        {
            l->lowresMvs[k][i] = (MV*)p; p += alignUp((size_t)n * 4, 64);
            l->lowresMvCosts[k][i] = (int32_t*)p; p += alignUp((size_t)n * 4, 64);
        }
    l->width = m_geom.width; l->lines = m_geom.lines; l->lumaStride = m_geom.stride; l->bframes = bf;
    l->slot = m_freeSlots.back();
    m_freeSlots.pop_back();
    return l;
}

/* x265's own `struct Lowres` behind this layer: the arrays are the caller's */
Lowres* Lookahead::adoptLowres(const Lowres& arrays)
{
    if (m_freeSlots.empty()) { snprintf(m_error, sizeof(m_error), "adoptLowres: no free frame slot"); return NULL; }
    Lowres* l = (Lowres*)calloc(1, sizeof(Lowres));
    if (!l) return NULL;
    *l = arrays;
    l->arena = NULL; l->arenaBytes = 0;
    l->width = m_geom.width; l->lines = m_geom.lines; l->lumaStride = m_geom.stride; l->bframes = m_param.bframes;
    l->ready = false; l->propagateStale = false;
    memset(l->mvVersion, 0, sizeof(l->mvVersion)); memset(l->devMvVersion, 0, sizeof(l->devMvVersion));
    memset(l->costStamp, 0, sizeof(l->costStamp)); memset(l->devCostStamp, 0, sizeof(l->devCostStamp));
    l->slot = m_freeSlots.back();
    m_freeSlots.pop_back();
    return l;
}

/* a frame is being re-initialised or released: nothing computed ahead from it may survive */
void Lookahead::forgetFrame(Lowres* l)
{
    std::map<int, Lowres*>::iterator it = m_byPoc.find(l->frameNum);
    if (it != m_byPoc.end() && it->second == l) m_byPoc.erase(it);
    for (size_t i = 0; i < m_spec.size();)
        if (m_spec[i].fenc == l || m_spec[i].ref0 == l || m_spec[i].ref1 == l) { m_spec[i] = m_spec.back(); m_spec.pop_back(); }
        else i++;
    l->ready = false;
    for (size_t i = 0; i < m_ctTouched.size();)
        if (m_ctTouched[i] == l) { m_ctTouched[i] = m_ctTouched.back(); m_ctTouched.pop_back(); }
        else i++;
}

void Lookahead::freeLowres(Lowres* l)
{
    if (!l) return;
    forgetFrame(l);
    /* copies into this frame's arrays may still be queued or in flight (plane copy-backs run behind the compute stream and
     * those of a pre-lookahead list are held back until the next estimate batch): let them land before the memory goes */
    if (m_ctx) x265cu_sync(m_ctx);
    m_freeSlots.push_back(l->slot);
    if (l->arena)
    {
        x265cu_host_unregister(l->arena);
        free(l->arena);
    }
    free(l);
}

/* Lowres::init, common/lowres.cpp:128-165 */
void Lookahead::lowresReset(Lowres& l, int poc)
{
    forgetFrame(&l);
    if (!m_ctOps.empty()) cuTreeRun(NULL, 0);   /* queued cuTree steps still read this frame's old arrays */
    memset(l.mvVersion, 0, sizeof(l.mvVersion));
    memset(l.devMvVersion, 0, sizeof(l.devMvVersion));
    memset(l.costStamp, 0, sizeof(l.costStamp));
    memset(l.devCostStamp, 0, sizeof(l.devCostStamp));
    l.frameNum = poc;
    memset(l.costEst, -1, sizeof(l.costEst));
    memset(l.weightedCostDelta, 0, sizeof(l.weightedCostDelta));
    memset(l.weightedRef, 0, sizeof(l.weightedRef));
    if (l.qpAqOffset && l.invQscaleFactor)
        memset(l.costEstAq, -1, sizeof(l.costEstAq));
    for (int y = 0; y < l.bframes + 2; y++)
        for (int x = 0; x < l.bframes + 2; x++)
            l.rowSatds[y][x][0] = -1;
    for (int i = 0; i < l.bframes + 1; i++)
    {
        l.lowresMvs[0][i][0].x = 0x7FFF;
        l.lowresMvs[1][i][0].x = 0x7FFF;
    }
    for (int i = 0; i < l.bframes + 2; i++)
        l.intraMbs[i] = 0;
}

bool Lookahead::lowresInit(Lowres& l, const void* luma, intptr_t stride, int poc, bool copyPlanesBack)
{
    lowresReset(l, poc);
    int r = x265cu_frame_init(m_ctx, l.slot, luma, stride, m_resident ? 1 : 0, (copyPlanesBack && !m_resident) ? l.buffer[0] : NULL);
    if (r) { snprintf(m_error, sizeof(m_error), "x265cu_frame_init: %s", x265cu_last_error(m_ctx)); return false; }
    return true;
}

/* aq-mode 1 (slicetype.cpp:193-207): qp_adj = strength * (log2(max(energy, 1)) - const), blocks independent of each other;
 * block rows [byFirst, byLast).  One compiled instance serves the whole-frame call and the row slices of the pool. */
__attribute__((noinline)) void Lookahead::aqMapRows(Lowres& l, const uint32_t* energy, const float* quantOffsets, int byFirst, int byLast)
{
    const Param& param = m_param;
    const int blocksX = (param.sourceWidth + 15) / 16;
    const double strength = param.aqStrength * 1.0397f;
    double qp_adj;
    int blockXY = byFirst * blocksX;
    for (int by = byFirst; by < byLast; by++)
        for (int bx = 0; bx < blocksX; bx++)
        {
            uint32_t e = energy[blockXY];
            qp_adj = strength * (log2((double)(e > 1 ? e : 1)) - (14.427f + 2 * (param.bitDepth - 8)));
            if (quantOffsets != NULL)
                qp_adj += quantOffsets[blockXY];
            l.qpAqOffset[blockXY] = qp_adj;
            l.qpCuTreeOffset[blockXY] = qp_adj;
            l.invQscaleFactor[blockXY] = exp2fix8(qp_adj);
            blockXY++;
        }
}

/* what calcAdaptiveQuantFrame does around the mapping: wp_sum / wp_ssd from the six frame sums (slicetype.cpp:138-160, 209-227) */
void Lookahead::aqFrameSums(Lowres& l, const uint64_t* sums)
{
    const Param& param = m_param;
    for (int i = 0; i < 3; i++) { l.wp_sum[i] = sums[i]; l.wp_ssd[i] = sums[3 + i]; }
    if (param.bEnableWeightedPred || param.bEnableWeightedBiPred)
    {
        int maxCol = ((param.sourceWidth + 8) >> 4) << 4;
        int maxRow = ((param.sourceHeight + 8) >> 4) << 4;
        int width[3] = { maxCol, maxCol >> 1, maxCol >> 1 };
        int height[3] = { maxRow, maxRow >> 1, maxRow >> 1 };
        for (int i = 0; i < 3; i++)
        {
            uint64_t sum = l.wp_sum[i], ssd = l.wp_ssd[i];
            l.wp_ssd[i] = ssd - (sum * sum + (width[i] * height[i]) / 2) / (width[i] * height[i]);
        }
    }
}

/* LookaheadTLD::calcAdaptiveQuantFrame, encoder/slicetype.cpp:95-228.
 * The per-block AC energy and the wp sums come from the GPU; the mapping below is the host float. */
bool Lookahead::calcAdaptiveQuantFrame(Lowres& l, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride,
                                       const uint32_t* preEnergy, const uint64_t* preSums, bool publish, const float* quantOffsets)
{
    const Param& param = m_param;
    int maxCol = param.sourceWidth, maxRow = param.sourceHeight;
    const int blocksX = (maxCol + 15) / 16, blocksY = (maxRow + 15) / 16;
    const int blockCount = m_cuCount;
    const bool bWeighted = param.bEnableWeightedPred || param.bEnableWeightedBiPred;      /* slicetype.cpp:138,211 */
    const bool needVar = !(param.aqMode == 0 || param.aqStrength == 0) || bWeighted;
    std::vector<uint32_t> energy((size_t)blocksX * blocksY);
    uint64_t sums[6] = { 0, 0, 0, 0, 0, 0 };
    if (needVar && preEnergy)
    {
        /* already measured by the fused x265cu_frame_init_var call of preLookahead() */
        memcpy(&energy[0], preEnergy, energy.size() * sizeof(uint32_t));
        memcpy(sums, preSums, sizeof(sums));
    }
    else if (needVar)
    {
        int r = x265cu_frame_var(m_ctx, y, yStride, u, v, cStride, m_resident ? 1 : 0, &energy[0], sums);
        if (r) { snprintf(m_error, sizeof(m_error), "x265cu_frame_var: %s", x265cu_last_error(m_ctx)); return false; }
    }
    for (int i = 0; i < 3; i++) { l.wp_sum[i] = sums[i]; l.wp_ssd[i] = sums[3 + i]; }

    double strength = 0.f;
    if (param.aqMode == 0 || param.aqStrength == 0)
    {
        if (param.aqMode && param.aqStrength == 0)
        {
            if (quantOffsets)
                for (int i = 0; i < blockCount; i++)
                {
                    l.qpCuTreeOffset[i] = l.qpAqOffset[i] = quantOffsets[i];
                    l.invQscaleFactor[i] = exp2fix8(l.qpCuTreeOffset[i]);
                }
            else
            {
                memset(l.qpCuTreeOffset, 0, blockCount * sizeof(double));
                memset(l.qpAqOffset, 0, blockCount * sizeof(double));
                for (int i = 0; i < blockCount; i++) l.invQscaleFactor[i] = 256;
            }
        }
    }
    else
    {
        int blockXY = 0;
        double avg_adj_pow2 = 0, avg_adj = 0, qp_adj = 0;
        double bias_strength = 0.f;
        if (param.aqMode == 2 || param.aqMode == 3)
        {
            double bit_depth_correction = 1.f / (1 << (2 * (param.bitDepth - 8)));
            l.frameVariance = 0;
            for (int by = 0; by < blocksY; by++)
            {
                uint64_t rowVariance = 0;
                for (int bx = 0; bx < blocksX; bx++)
                {
                    uint32_t e = energy[blockXY];
                    l.blockVariance[blockXY] = e;
                    rowVariance += e;
                    qp_adj = pow(e * bit_depth_correction + 1, 0.1);
                    l.qpCuTreeOffset[blockXY] = qp_adj;
                    avg_adj += qp_adj;
                    avg_adj_pow2 += qp_adj * qp_adj;
                    blockXY++;
                }
                l.frameVariance += (rowVariance / maxCol);
            }
            l.frameVariance /= maxRow;
            avg_adj /= blockCount;
            avg_adj_pow2 /= blockCount;
            strength = param.aqStrength * avg_adj;
            avg_adj = avg_adj - 0.5f * (avg_adj_pow2 - (11.f)) / avg_adj;
            bias_strength = param.aqStrength;
        }
        else
            strength = param.aqStrength * 1.0397f;

        blockXY = 0;
        if (param.aqMode == 1)
            aqMapRows(l, &energy[0], quantOffsets, 0, blocksY);
        else
        for (int by = 0; by < blocksY; by++)
            for (int bx = 0; bx < blocksX; bx++)
            {
                if (param.aqMode == 3)
                {
                    qp_adj = l.qpCuTreeOffset[blockXY];
                    qp_adj = strength * (qp_adj - avg_adj) + bias_strength * (1.f - 11.f / (qp_adj * qp_adj));
                }
                else if (param.aqMode == 2)
                {
                    qp_adj = l.qpCuTreeOffset[blockXY];
                    qp_adj = strength * (qp_adj - avg_adj);
                }
                else
                {
                    uint32_t e = energy[blockXY];
                    qp_adj = strength * (log2((double)(e > 1 ? e : 1)) - (14.427f + 2 * (param.bitDepth - 8)));
                }
                if (quantOffsets != NULL)
                    qp_adj += quantOffsets[blockXY];
                l.qpAqOffset[blockXY] = qp_adj;
                l.qpCuTreeOffset[blockXY] = qp_adj;
                l.invQscaleFactor[blockXY] = exp2fix8(qp_adj);
                blockXY++;
            }
    }
    if (bWeighted)
    {
        maxCol = ((maxCol + 8) >> 4) << 4;
        maxRow = ((maxRow + 8) >> 4) << 4;
        int width[3] = { maxCol, maxCol >> 1, maxCol >> 1 };
        int height[3] = { maxRow, maxRow >> 1, maxRow >> 1 };
        for (int i = 0; i < 3; i++)
        {
            uint64_t sum = l.wp_sum[i], ssd = l.wp_ssd[i];
            l.wp_ssd[i] = ssd - (sum * sum + (width[i] * height[i]) / 2) / (width[i] * height[i]);
        }
    }
    if (!publish) return true;     /* x265cu_pre_lookahead_batch uploads the array this call returns to it */
    int r = x265cu_frame_set_invqscale(m_ctx, l.slot, l.invQscaleFactor);
    if (r) { snprintf(m_error, sizeof(m_error), "x265cu_frame_set_invqscale: %s", x265cu_last_error(m_ctx)); return false; }
    return true;
}

/* LookaheadTLD::lowresIntraEstimate, encoder/slicetype.cpp:230-336 */
bool Lookahead::lowresIntraEstimate(Lowres& l)
{
    x265cu_intra_out o;
    o.intraCost = l.intraCost; o.intraMode = l.intraMode;
    o.lowresCosts = l.lowresCosts[0][0]; o.rowSatds = l.rowSatds[0][0];
    if (m_resident)
    {
        /* device-resident mode: arrays stay in the HBM mirrors, only the two sums come back */
        o.intraCost = NULL; o.intraMode = NULL; o.lowresCosts = NULL; o.rowSatds = NULL;
        l.rowSatds[0][0][0] = 0;
    }
    int r = x265cu_intra(m_ctx, l.slot, &o);
    if (r) { snprintf(m_error, sizeof(m_error), "x265cu_intra: %s", x265cu_last_error(m_ctx)); return false; }
    l.costEst[0][0] = o.sums[0];
    l.costEstAq[0][0] = o.sums[1];
    return true;
}

/* PreLookaheadGroup::processTasks for one frame, encoder/slicetype.cpp:831-856 */
bool Lookahead::preLookahead(Lowres& l, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride, int poc, bool copyPlanesBack,
                             const float* quantOffsets)
{
    const bool needVar = m_bAdaptiveQuant && (!(m_param.aqMode == 0 || m_param.aqStrength == 0) || m_param.bEnableWeightedPred || m_param.bEnableWeightedBiPred);
    if (needVar)
    {
        /* Lowres::init and acEnergyCu share one upload of the picture */
        lowresReset(l, poc);
        const int blocks = ((m_param.sourceWidth + 15) / 16) * ((m_param.sourceHeight + 15) / 16);
        std::vector<uint32_t> energy((size_t)blocks);
        uint64_t sums[6];
        int r = x265cu_frame_init_var(m_ctx, l.slot, y, yStride, u, v, cStride, m_resident ? 1 : 0,
                                      (copyPlanesBack && !m_resident) ? l.buffer[0] : NULL, &energy[0], sums);
        if (r) { snprintf(m_error, sizeof(m_error), "x265cu_frame_init_var: %s", x265cu_last_error(m_ctx)); return false; }
        if (!calcAdaptiveQuantFrame(l, y, yStride, u, v, cStride, &energy[0], sums, true, quantOffsets)) return false;
    }
    else
    {
        if (!lowresInit(l, y, yStride, poc, copyPlanesBack)) return false;
        if (m_bAdaptiveQuant)
        {
            if (!calcAdaptiveQuantFrame(l, y, yStride, u, v, cStride, NULL, NULL, true, quantOffsets)) return false;
        }
        else
            x265cu_frame_set_invqscale(m_ctx, l.slot, NULL);
    }
    if (!lowresIntraEstimate(l)) return false;
    l.ready = true;
    m_byPoc[poc] = &l;
    m_newestReady++;
    return true;
}

/* Lookahead::addPicture, encoder/slicetype.cpp:633-650: start the picture's upload when it enters the input queue */
bool Lookahead::addPicture(Lowres& l, const PictureIn& pic)
{
    if (m_resident || !pic.y || !pic.u || !pic.v) return true;      /* nothing to upload ahead */
    if (getenv("X265CU_PREFETCH") && atoi(getenv("X265CU_PREFETCH")) == 0) return true;
    int r = x265cu_frame_upload(m_ctx, l.slot, pic.y, pic.yStride, pic.u, pic.v, pic.cStride);
    if (r) { snprintf(m_error, sizeof(m_error), "x265cu_frame_upload: %s", x265cu_last_error(m_ctx)); return false; }
    return true;
}

/* PreLookaheadGroup::processTasks for its whole list m_preframes[0..m_jobTotal) (slicetype.cpp:831-856): the reference
 * spreads the frames of the list over worker threads; here the uploads and kernels of all of them are enqueued back to
 * back and the host waits once per stage (lowres + variance, then intra) instead of twice per frame. */
bool Lookahead::preLookaheadBatch(int n, Lowres** ls, const PictureIn* pics, bool copyPlanesBack)
{
    const bool needVar = m_bAdaptiveQuant && (!(m_param.aqMode == 0 || m_param.aqStrength == 0) || m_param.bEnableWeightedPred || m_param.bEnableWeightedBiPred);
    if (!needVar || n < 2)
    {
        for (int i = 0; i < n; i++)
            if (!preLookahead(*ls[i], pics[i].y, pics[i].yStride, pics[i].u, pics[i].v, pics[i].cStride, pics[i].poc, copyPlanesBack, pics[i].quantOffsets)) return false;
        return true;
    }
    const int blocks = ((m_param.sourceWidth + 15) / 16) * ((m_param.sourceHeight + 15) / 16);
    std::vector<uint32_t> energy((size_t)blocks * n);
    std::vector<uint64_t> sums((size_t)6 * n);
    std::vector<x265cu_frame_in> items((size_t)n);
    for (int i = 0; i < n; i++)
    {
        lowresReset(*ls[i], pics[i].poc);
        x265cu_frame_in& f = items[i];
        f.slot = ls[i]->slot;
        f.y = pics[i].y; f.yStride = pics[i].yStride; f.u = pics[i].u; f.v = pics[i].v; f.cStride = pics[i].cStride;
        f.planesAreDevice = m_resident ? 1 : 0;
        f.planesOut = (copyPlanesBack && !m_resident) ? ls[i]->buffer[0] : NULL;
        f.energy = &energy[(size_t)blocks * i];
        f.sums = &sums[(size_t)6 * i];
    }
    std::vector<x265cu_intra_out> outs((size_t)n);
    for (int i = 0; i < n; i++)
    {
        Lowres& l = *ls[i];
        x265cu_intra_out& o = outs[i];
        o.intraCost = l.intraCost; o.intraMode = l.intraMode;
        o.lowresCosts = l.lowresCosts[0][0]; o.rowSatds = l.rowSatds[0][0];
        if (m_resident)
        {
            o.intraCost = NULL; o.intraMode = NULL; o.lowresCosts = NULL; o.rowSatds = NULL;
            l.rowSatds[0][0][0] = 0;
        }
    }
    /* one pipelined call: while the pictures of the later frames are still crossing PCIe, the float AQ mapping of
     * frame i runs here (callback, same code and flags as ever) and its intra estimate starts on the GPU */
    struct AqCtx { Lookahead* la; Lowres** ls; const PictureIn* pics; const x265cu_frame_in* items; bool ok; } actx = { this, ls, pics, &items[0], true };
    struct Aq
    {
        static void one(AqCtx* a, int i)
        {
            if (!a->la->calcAdaptiveQuantFrame(*a->ls[i], a->pics[i].y, a->pics[i].yStride, a->pics[i].u, a->pics[i].v, a->pics[i].cStride,
                                               a->items[i].energy, a->items[i].sums, false, a->pics[i].quantOffsets))
                a->ok = false;
        }
        static void run(void* user, int first, int count, const int32_t** invQ)
        {
            AqCtx* a = (AqCtx*)user;
            /* the mapping is per frame and independent: the reference runs the frames of the list on different workers,
             * so do we for a run of several frames (same code, same flags, same results); in aq-mode 1 the block rows of a
             * frame are independent too and are spread as well */
            Lookahead* la = a->la;
            const Param& p = la->m_param;
            HostPool& pool = HostPool::get();
            if (p.aqMode == 1 && p.aqStrength != 0)
            {
                const int blocksY = (p.sourceHeight + 15) / 16;
                int slices = (pool.threads() + count - 1) / count;
                if (slices > blocksY) slices = blocksY;
                if (slices < 1) slices = 1;
                pool.run(count * slices, [a, la, first, slices, blocksY](int k)
                {
                    const int i = first + k / slices, sl = k % slices;
                    const int r0 = (int)((int64_t)blocksY * sl / slices), r1 = (int)((int64_t)blocksY * (sl + 1) / slices);
                    if (sl == 0) la->aqFrameSums(*a->ls[i], a->items[i].sums);
                    la->aqMapRows(*a->ls[i], a->items[i].energy, a->pics[i].quantOffsets, r0, r1);
                });
            }
            else
                pool.run(count, [a, first](int k) { one(a, first + k); });
            for (int i = 0; i < count; i++) invQ[i] = a->ls[first + i]->invQscaleFactor;
        }
    };
    if (getenv("X265CU_PRE_PIPELINE") && atoi(getenv("X265CU_PRE_PIPELINE")) == 0)
    {
        /* the same work as three calls with a wait after each (tests: both forms must give identical results) */
        int r3 = x265cu_frame_init_var_batch(m_ctx, n, &items[0]);
        if (r3) { snprintf(m_error, sizeof(m_error), "x265cu_frame_init_var_batch: %s", x265cu_last_error(m_ctx)); return false; }
        std::vector<int> slots((size_t)n);
        for (int i = 0; i < n; i++)
        {
            if (!calcAdaptiveQuantFrame(*ls[i], pics[i].y, pics[i].yStride, pics[i].u, pics[i].v, pics[i].cStride, items[i].energy, items[i].sums, true, pics[i].quantOffsets)) return false;
            slots[i] = ls[i]->slot;
        }
        r3 = x265cu_intra_batch(m_ctx, n, &slots[0], &outs[0]);
        if (r3) { snprintf(m_error, sizeof(m_error), "x265cu_intra_batch: %s", x265cu_last_error(m_ctx)); return false; }
    }
    else
    {
        int r = x265cu_pre_lookahead_batch(m_ctx, n, &items[0], Aq::run, &actx, &outs[0]);
        if (r) { snprintf(m_error, sizeof(m_error), "x265cu_pre_lookahead_batch: %s", x265cu_last_error(m_ctx)); return false; }
        if (!actx.ok) return false;
    }
    for (int i = 0; i < n; i++)
    {
        Lowres& l = *ls[i];
        l.costEst[0][0] = outs[i].sums[0];
        l.costEstAq[0][0] = outs[i].sums[1];
        l.ready = true;
        m_byPoc[pics[i].poc] = &l;
        m_newestReady++;
    }
    return true;
}

/* ------------------------------------------------------------------------------------------
 * CostEstimateGroup, encoder/slicetype.cpp:1899-2066
 * ---------------------------------------------------------------------------------------- */
void CostEstimateGroup::add(int p0, int p1, int b)
{
    m_batchMode = true;
    Estimate& e = m_estimates[m_jobTotal++];
    e.p0 = p0; e.p1 = p1; e.b = b;
    if (m_jobTotal == MAX_BATCH_SIZE)
        finishBatch();
}

/* the still unknown estimates of ep[first..) shifted by `shift` frames, as look-ahead requests (non-batch estimates:
 * cooperative slices).  valid = every frame they need exists and went through preLookahead(). */
void CostEstimateGroup::predictFrom(const std::vector<Lookahead::Request>& ep, size_t first, int shift, Lowres* skipFenc, int skipD0, int skipD1,
                                    const std::vector<EstReq>& already, std::vector<EstReq>& out, bool& valid)
{
    Lookahead& la = m_lookahead;
    valid = true;
    for (size_t i = first; i < ep.size(); i++)
    {
        const Lookahead::Request& q = ep[i];
        std::map<int, Lowres*>::iterator ib = la.m_byPoc.find(q.b + shift), i0 = la.m_byPoc.find(q.p0 + shift), i1 = la.m_byPoc.find(q.p1 + shift);
        if (ib == la.m_byPoc.end() || i0 == la.m_byPoc.end() || i1 == la.m_byPoc.end()) { valid = false; return; }
        Lowres *qb = ib->second, *q0 = i0->second, *q1 = i1->second;
        const int e0 = q.b - q.p0, e1 = q.p1 - q.b;
        if (!qb->ready || !q0->ready || !q1->ready) { valid = false; return; }
        if (e0 < 1 || e0 > la.m_param.bframes + 1 || e1 < 0 || e1 > la.m_param.bframes + 1) continue;
        if (qb->costEst[e0][e1] >= 0 && qb->rowSatds[e0][e1][0] != -1) continue;      /* already known */
        bool dup = qb == skipFenc && e0 == skipD0 && e1 == skipD1;
        for (size_t c = 0; c < out.size(); c++)
            dup = dup || (out[c].fenc == qb && out[c].d0 == e0 && out[c].d1 == e1);
        for (size_t c = 0; c < already.size(); c++)
            dup = dup || (already[c].fenc == qb && already[c].d0 == e0 && already[c].d1 == e1);
        for (size_t c = 0; c < la.m_spec.size(); c++)      /* still pending from an earlier prediction */
            dup = dup || (la.m_spec[c].fenc == qb && la.m_spec[c].d0 == e0 && la.m_spec[c].d1 == e1);
        if (dup) continue;
        EstReq r = { qb, q0, q1, e0, e1, true, true };
        out.push_back(r);
    }
}

/* Rule-based part of the look-ahead cache, for requests the history cannot foresee (the first slicetypeDecide of a stream,
 * a new kind of request).  The non-batch estimates of a b-adapt 2 analysis are those of the trellis: slicetypePath
 * (slicetype.cpp:1565-1592) tries, for every path length N, the suffixes "P", "BP", "BBP", ... and slicetypePathCost
 * (:1594-1639) asks for the P estimate of the last segment (N - d -> N) and for its B estimates (b-pyramid: the middle frame
 * between the two P frames, then the two halves).  The batches of slicetypeAnalyse (:1231-1297) only cover b >= 2 and L1
 * searches at the distance of an L0 search, below the last frame: what is left are the estimates next to the first frame of
 * the analysis and every segment that ends at its LAST frame -- 13 (bframes 4) to 27 (bframes 8) one-by-one requests in the first
 * decision of a stream.  They are listed here in the trellis's own order (the first estimate to find a field's sentinel
 * searches it, and the vectors of a field depend on the kind of estimate that did, :2146-2160) and ride in the launch of the
 * request that was not foreseen.  Nothing here can change a result: an entry is handed out only to exactly the request it
 * computed (takeAhead); a wrong guess is dropped at the next batch. */
void CostEstimateGroup::predictTrellis(const Lookahead::Request& rq, Lowres* skipFenc, int skipD0, int skipD1, const std::vector<EstReq>& already,
                                       std::vector<EstReq>& out)
{
    Lookahead& la = m_lookahead;
    const int maxD = la.m_param.bframes + 1;
    int first = rq.p0, last = rq.p1;
    if (la.m_batchFirst >= 0 && la.m_batchFirst <= rq.p0 && rq.p0 - la.m_batchFirst <= la.m_param.lookaheadDepth + maxD) first = la.m_batchFirst;
    if (la.m_batchLast >= 0 && la.m_batchLast + 1 > last && la.m_batchLast + 1 - rq.p1 <= la.m_param.lookaheadDepth + maxD) last = la.m_batchLast + 1;
    std::vector<Lookahead::Request> ep;
    for (int N = first + 1; N <= last; N++)
    {
        if (la.m_byPoc.find(N) == la.m_byPoc.end()) break;          /* a stream's frames arrive in order */
        for (int d = 1; d <= maxD && N - d >= first; d++)
        {
            const int c = N - d;
            if (la.m_byPoc.find(c) == la.m_byPoc.end()) break;
            bool all = true;
            for (int b = c + 1; b < N; b++) all = all && la.m_byPoc.find(b) != la.m_byPoc.end();
            if (!all) break;
            Lookahead::Request q = { c, N, N };
            ep.push_back(q);
            if (la.m_param.bBPyramid && d > 2)
            {
                const int middle = c + d / 2;
                Lookahead::Request m = { c, middle, N };
                ep.push_back(m);
                for (int b = c + 1; b < middle; b++) { Lookahead::Request r = { c, b, middle }; ep.push_back(r); }
                for (int b = middle + 1; b < N; b++) { Lookahead::Request r = { middle, b, N }; ep.push_back(r); }
            }
            else
                for (int b = c + 1; b < N; b++) { Lookahead::Request r = { c, b, N }; ep.push_back(r); }
            /* slicetypeDecide's own frame costs for the mini-GOP it settles on (slicetype.cpp:992-1035) use the B-ref it
             * inserts at list[bframes / 2]; for an even number of B frames that is not the trellis's middle frame */
            const int nb = d - 1, bref = c + nb / 2 + 1;
            if (la.m_trellisBref && la.m_param.bBPyramid && nb > 1 && bref != c + d / 2)
            {
                Lookahead::Request m = { c, bref, N };
                ep.push_back(m);
                for (int b = c + 1; b < bref; b++) { Lookahead::Request r = { c, b, bref }; ep.push_back(r); }
                for (int b = bref + 1; b < N; b++) { Lookahead::Request r = { bref, b, N }; ep.push_back(r); }
            }
        }
    }
    bool valid = true;
    std::vector<EstReq> more;
    predictFrom(ep, 0, 0, skipFenc, skipD0, skipD1, already, more, valid);
    if (!valid) return;
    size_t aheadAlready = 0;
    for (size_t k = 0; k < already.size(); k++) aheadAlready += already[k].ahead ? 1 : 0;
    const size_t room = aheadAlready < 192 ? 192 - aheadAlready : 0;   /* bounds the side buffers of one call */
    if (more.size() > room) more.resize(room);
    out.swap(more);
}

bool CostEstimateGroup::finishBatch()
{
    Lookahead& la = m_lookahead;
    std::vector<EstReq> reqs((size_t)m_jobTotal);
    for (int i = 0; i < m_jobTotal; i++)
    {
        const Estimate& e = m_estimates[i];
        EstReq r = { m_frames[e.b], m_frames[e.p0], m_frames[e.p1], e.b - e.p0, e.p1 - e.b, false, false };
        reqs[i] = r;
        const int lo = m_frames[e.p0]->frameNum, hi = m_frames[e.p1]->frameNum;
        if (i == 0 || lo < la.m_batchFirst) la.m_batchFirst = lo;
        if (i == 0 || hi > la.m_batchLast) la.m_batchLast = hi;
    }
    if (m_jobTotal && !la.m_episode.empty())
    {
        /* a batch closes the run of one-by-one estimates: what was computed ahead and not asked for is dropped,
         * the run becomes the pattern the next one is predicted from */
        la.m_spec.clear();
        la.m_history.push_back(la.m_episode);
        la.m_historyNewest.push_back(la.m_episodeNewest);
        if (la.m_history.size() > 3) { la.m_history.erase(la.m_history.begin()); la.m_historyNewest.erase(la.m_historyNewest.begin()); }
        la.m_episode.clear();
        /* ... and the next run is about to begin: it will most likely repeat the last one, shifted by as many frames
         * as have arrived since that one began.  Its estimates ride along in this batch's launch. */
        if (la.m_lookAhead)
        {
            const int shift = la.m_newestReady - la.m_historyNewest.back();
            std::vector<EstReq> ahead;
            bool valid = false;
            if (shift > 0) predictFrom(la.m_history.back(), 0, shift, NULL, 0, 0, reqs, ahead, valid);
            if (valid) reqs.insert(reqs.end(), ahead.begin(), ahead.end());
            if (getenv("X265CU_LOOKAHEAD_DEBUG"))
                fprintf(stderr, "finishBatch: %d jobs, shift %d (newest %d, then %d), pattern of %d, valid %d, %d ahead\n", m_jobTotal, shift,
                        la.m_newestReady, la.m_historyNewest.back(), (int)la.m_history.back().size(), (int)valid, (int)ahead.size());
        }
    }
    if (la.m_lookAhead && la.m_trellisAtBatch && m_jobTotal && la.m_batchLast >= 0)
    {
        /* ... and what the history does not cover of the trellis that follows this batch (its segments up to the frame
         * after the batch's last one: the analysis's last frame) rides along too */
        const Lookahead::Request rq = { la.m_batchLast, la.m_batchLast + 1, la.m_batchLast + 1 };
        std::vector<EstReq> all, more;
        predictTrellis(rq, NULL, 0, 0, reqs, all);
        /* only what the frame-cost batch that follows cannot contain: estimates that touch the frame after the batch's
         * last one (the batches stop below it, slicetype.cpp:1235,1263,1276); everything else it is about to compute */
        for (size_t k = 0; k < all.size(); k++)
            if (all[k].fenc->frameNum == rq.p1 || all[k].ref1->frameNum == rq.p1) more.push_back(all[k]);
        reqs.insert(reqs.end(), more.begin(), more.end());
        if (getenv("X265CU_LOOKAHEAD_DEBUG") && !more.empty())
            fprintf(stderr, "finishBatch: %d more ahead by the trellis rule\n", (int)more.size());
    }
    bool ok = reqs.empty() ? true : runEstimates(&reqs[0], (int)reqs.size());
    m_jobTotal = 0;
    return ok;
}

/* hand out an estimate that was computed ahead, if the request is exactly that computation */
bool CostEstimateGroup::takeAhead(Lowres* fenc, Lowres* ref0, Lowres* ref1, int d0, int d1)
{
    Lookahead& la = m_lookahead;
    const int n = la.m_cuCount, rows = la.m_8x8Height;
    for (size_t i = 0; i < la.m_spec.size(); i++)
    {
        SpecEstimate& e = la.m_spec[i];
        if (e.fenc != fenc || e.ref0 != ref0 || e.ref1 != ref1 || e.d0 != d0 || e.d1 != d1) continue;
        if (e.fencNum != fenc->frameNum || e.ref0Num != ref0->frameNum || e.ref1Num != ref1->frameNum) continue;
        const int s0 = fenc->lowresMvs[0][d0 - 1][0].x == 0x7FFF;
        const int s1 = d1 > 0 && fenc->lowresMvs[1][d1 - 1][0].x == 0x7FFF;
        if (s0 != e.doSearch[0] || s1 != e.doSearch[1]) continue;
        /* the MV fields it consumed must be the ones that are official now */
        if (!s0 && fenc->mvVersion[0][d0 - 1] != e.usedVersion[0]) continue;
        if (d1 > 0 && !s1 && fenc->mvVersion[1][d1 - 1] != e.usedVersion[1]) continue;
        if (la.m_resident)
        {
            /* the arrays only exist as device mirrors: they must still be what THIS computation wrote (a later estimate
             * computed ahead for the same slot of the tables may have overwritten them) */
            if (fenc->devCostStamp[d0][d1] != e.costStamp) continue;
            if (e.doSearch[0] && fenc->devMvVersion[0][d0 - 1] != e.newVersion[0]) continue;
            if (e.doSearch[1] && fenc->devMvVersion[1][d1 - 1] != e.newVersion[1]) continue;
        }

        if (!la.m_resident)
        {
            for (int l = 0; l < 2; l++)
                if (e.doSearch[l])
                {
                    const int d = l ? d1 : d0;
                    memcpy(fenc->lowresMvs[l][d - 1], &e.mvs[l][0], (size_t)n * 4);
                    memcpy(fenc->lowresMvCosts[l][d - 1], &e.mvCosts[l][0], (size_t)n * 4);
                }
            memcpy(fenc->lowresCosts[d0][d1], &e.lowresCosts[0], (size_t)n * 2);
            memcpy(fenc->rowSatds[d0][d1], &e.rowSatds[0], (size_t)rows * 4);
        }
        else
        {
            if (e.doSearch[0]) fenc->lowresMvs[0][d0 - 1][0].x = 0;
            if (e.doSearch[1]) fenc->lowresMvs[1][d1 - 1][0].x = 0;
            fenc->rowSatds[d0][d1][0] = 0;
        }
        for (int l = 0; l < 2; l++)
            if (e.doSearch[l]) fenc->mvVersion[l][(l ? d1 : d0) - 1] = e.newVersion[l];
        fenc->costStamp[d0][d1] = e.costStamp;
        fenc->costEst[d0][d1] = e.res.costEst;
        fenc->costEstRaw[d0][d1] = e.res.costEstRaw;
        fenc->costEstAq[d0][d1] = e.res.costEstAq;
        if (d1 == 0) fenc->intraMbs[d0] += e.res.intraMbs;
        fenc->weightedRef[d0].present = 0;
        if (e.wref.present) { fenc->weightedRef[d0] = e.wref; fenc->weightedCostDelta[d0] = e.wdelta; }
        la.m_spec[i] = la.m_spec.back();
        la.m_spec.pop_back();
        la.m_specStats[1]++;
        return true;
    }
    return false;
}

int64_t CostEstimateGroup::singleCost(int p0, int p1, int b, bool intraPenalty)
{
    Lookahead& la = m_lookahead;
    Lowres* fenc = m_frames[b];
    const int d0 = b - p0, d1 = p1 - b;
    int64_t score;
    if (fenc->costEst[d0][d1] >= 0 && fenc->rowSatds[d0][d1][0] != -1)
        score = fenc->costEst[d0][d1];
    else
    {
        la.m_specStats[3]++;
        const Lookahead::Request rq = { m_frames[p0]->frameNum, fenc->frameNum, m_frames[p1]->frameNum };
        if (!takeAhead(fenc, m_frames[p0], m_frames[p1], d0, d1))
        {
            std::vector<EstReq> reqs;
            EstReq r0 = { fenc, m_frames[p0], m_frames[p1], d0, d1, false, true };
            reqs.push_back(r0);
            /* The request was not foreseen: look for the same kind of request (same distances) in the recent history,
             * shifted in time, and take the requests that followed it then as the ones about to be asked for now.
             * Among the alignments whose predicted frames all exist, the one predicting most wins. */
            if (la.m_lookAhead)
            {
                std::vector<EstReq> best;
                for (int h = (int)la.m_history.size(); h >= 0; h--)
                {
                    const std::vector<Lookahead::Request>& ep = h == (int)la.m_history.size() ? la.m_episode : la.m_history[h];
                    for (size_t k = 0; k < ep.size(); k++)
                    {
                        const int shift = rq.b - ep[k].b;
                        if (shift == 0 || ep[k].b - ep[k].p0 != d0 || ep[k].p1 - ep[k].b != d1) continue;
                        std::vector<EstReq> cand;
                        bool valid = true;
                        predictFrom(ep, k + 1, shift, fenc, d0, d1, std::vector<EstReq>(), cand, valid);
                        if (valid && cand.size() > best.size()) best.swap(cand);
                    }
                }
                reqs.insert(reqs.end(), best.begin(), best.end());
                if (la.m_trellisAhead)
                {
                    std::vector<EstReq> more;
                    predictTrellis(rq, fenc, d0, d1, reqs, more);
                    reqs.insert(reqs.end(), more.begin(), more.end());
                }
            }
            if (getenv("X265CU_LOOKAHEAD_DEBUG"))
                fprintf(stderr, "singleCost (%d,%d,%d): on demand, %d ahead, episode %d prev %d\n", rq.p0, rq.b, rq.p1, (int)reqs.size() - 1, (int)la.m_episode.size(), (int)la.m_history.size());
            if (!runEstimates(&reqs[0], (int)reqs.size())) return -1;
            la.m_specStats[2]++;
        }
        else if (getenv("X265CU_LOOKAHEAD_DEBUG"))
            fprintf(stderr, "singleCost (%d,%d,%d): handed out\n", rq.p0, rq.b, rq.p1);
        if (la.m_episode.empty()) la.m_episodeNewest = la.m_newestReady;
        la.m_episode.push_back(rq);
        score = fenc->costEst[d0][d1];
    }
    if (intraPenalty)
        score += score * fenc->intraMbs[d0] / (m_lookahead.ncu() * 8);
    return score;
}

namespace {
/* first half of LookaheadTLD::weightsAnalyse (slicetype.cpp:416-455): the float guesses */
struct WeightGuess { bool skip; int minscale, mindenom, curScale, curOffset; };

WeightGuess weightGuess(const Lowres& fenc, const Lowres& ref, int depth)
{
    static const float epsilon = 1.f / 128.f;
    WeightGuess g;
    g.skip = true; g.minscale = g.mindenom = g.curScale = g.curOffset = 0;
    float guessScale, fencMean, refMean;
    if (fenc.wp_ssd[0] && ref.wp_ssd[0])
        guessScale = sqrtf((float)fenc.wp_ssd[0] / ref.wp_ssd[0]);
    else
        guessScale = 1.0f;
    fencMean = (float)fenc.wp_sum[0] / (fenc.lines * fenc.width) / (1 << (depth - 8));
    refMean = (float)ref.wp_sum[0] / (fenc.lines * fenc.width) / (1 << (depth - 8));
    if (fabsf(refMean - fencMean) < 0.5f && fabsf(1.f - guessScale) < epsilon)
        return g;
    /* WeightParam::setFromWeightAndOffset((int)(guessScale * 128 + 0.5f), 0, 7, true), slice.h:292-305 */
    int denom = 7, w = (int)(guessScale * 128 + 0.5f);
    while (denom > 0 && w > 127) { denom--; w >>= 1; }
    w = imin(w, 127);
    g.mindenom = denom; g.minscale = w;
    int curScale = w;
    int curOffset = (int)(fencMean - refMean * curScale / (1 << denom) + 0.5f);
    if (curOffset < -128 || curOffset > 127)
    {
        curOffset = iclip(-128, 127, curOffset);
        curScale = (int)((1 << denom) * (fencMean - curOffset) / refMean + 0.5f);
        curScale = iclip(0, 127, curScale);
    }
    g.curScale = curScale; g.curOffset = curOffset;
    g.skip = false;
    return g;
}
} // namespace

/* est[0..n): estimates to compute in ONE x265cu_estimate_batch call.  Entries flagged `ahead` are computed into the
 * look-ahead cache instead of the frames' arrays; they may consume MV fields that an earlier entry of the same
 * call produces (the cost kernel of a batch runs after all of its searches). */
bool CostEstimateGroup::runEstimates(const EstReq* est, int n)
{
    Lookahead& la = m_lookahead;
    const Param& param = la.m_param;
    const int nCU = la.m_cuCount, rows = la.m_8x8Height;
    std::vector<x265cu_job> jobs;
    std::vector<int> jobOf;                 /* index into est[] */
    std::vector<WeightGuess> guesses;
    std::vector<x265cu_weight_item> witems;
    std::vector<int> wjob;                  /* job index of each pair of weight items */
    std::vector<WeightParam> wref;          /* per job: the weight decision for the L0 search */
    std::vector<double> wdelta;
    std::vector<int> specOf;                /* per job: index into la.m_spec, or -1 */
    /* MV fields this call produces: (frame, list, distance) -> version */
    struct Produced { Lowres* f; int l, d; uint64_t ver; };
    std::vector<Produced> produced;
    jobs.reserve((size_t)n);
    const size_t specBase = la.m_spec.size();
    la.m_spec.reserve(specBase + (size_t)n);   /* destination pointers into the entries must stay valid */

    for (int i = 0; i < n; i++)
    {
        Lowres *fenc = est[i].fenc, *ref0 = est[i].ref0, *ref1 = est[i].ref1;
        const int d0 = est[i].d0, d1 = est[i].d1;
        if (fenc->costEst[d0][d1] >= 0 && fenc->rowSatds[d0][d1][0] != -1)
            continue;                       /* cached (estimateFrameCost :1982) */
        x265cu_job j;
        memset(&j, 0, sizeof(j));
        j.fenc = fenc->slot; j.ref0 = ref0->slot; j.ref1 = ref1->slot;
        j.d0 = d0; j.d1 = d1;
        j.doSearch[0] = fenc->lowresMvs[0][d0 - 1][0].x == 0x7FFF;
        j.doSearch[1] = d1 > 0 && fenc->lowresMvs[1][d1 - 1][0].x == 0x7FFF;
        j.sliced = est[i].sliced;
        uint64_t usedVersion[2] = { fenc->mvVersion[0][d0 - 1], d1 > 0 ? fenc->mvVersion[1][d1 - 1] : 0 };
        uint64_t newVersion[2] = { 0, 0 };
        bool usable = true;
        for (int l = 0; l < 2; l++)
        {
            const int d = l ? d1 : d0;
            if (!j.doSearch[l]) continue;
            /* a field an earlier estimate of this call is already searching: consume that one (only look-ahead entries
             * can meet this; the reference's own batches are independent) */
            for (size_t k = 0; k < produced.size(); k++)
                if (produced[k].f == fenc && produced[k].l == l && produced[k].d == d)
                {
                    if (!est[i].ahead) usable = false;
                    j.doSearch[l] = 0;
                    usedVersion[l] = produced[k].ver;
                }
            if (j.doSearch[l])
            {
                newVersion[l] = ++la.m_versionCounter;
                Produced pr = { fenc, l, d, newVersion[l] };
                produced.push_back(pr);
                fenc->devMvVersion[l][d - 1] = newVersion[l];     /* this launch overwrites the device mirror of the field */
            }
        }
        /* lists this estimate does not search are CONSUMED from the device mirror of the field: that must be the official
         * field (an estimate computed ahead and never handed out may have searched the field again and left another one
         * there -- the vectors of a field depend on the kind of estimate that searched it, slicetype.cpp:2146-2160) */
        for (int l = 0; l < 2; l++)
        {
            const int d = l ? d1 : d0;
            if (j.doSearch[l] || (l && d1 == 0)) continue;
            bool fromThisCall = false;
            for (size_t k = 0; k < produced.size(); k++)
                if (produced[k].f == fenc && produced[k].l == l && produced[k].d == d) fromThisCall = true;
            if (fromThisCall) continue;
            if (fenc->devMvVersion[l][d - 1] != fenc->mvVersion[l][d - 1])
            {
                if (!la.m_resident)
                {
                    if (!la.m_ctOps.empty() && !la.cuTreeRun(NULL, 0)) return false;
                    if (x265cu_frame_set_array(la.m_ctx, fenc->slot, 6, l, d, fenc->lowresMvs[l][d - 1]))
                    {
                        snprintf(la.m_error, sizeof(la.m_error), "x265cu_frame_set_array: %s", x265cu_last_error(la.m_ctx));
                        return false;
                    }
                    fenc->devMvVersion[l][d - 1] = fenc->mvVersion[l][d - 1];
                }
                la.m_ctStats[2]++;
            }
            usedVersion[l] = fenc->devMvVersion[l][d - 1];     /* what the kernel will really read */
        }
        const uint64_t costStamp = ++la.m_versionCounter;
        fenc->devCostStamp[d0][d1] = costStamp;                   /* ... and of lowresCosts[d0][d1] */
        if (!usable) { snprintf(la.m_error, sizeof(la.m_error), "runEstimates: dependent estimates in one batch"); return false; }
        WeightParam noWeight = { 0, 0, 0, 0 };
        if (param.bEnableWeightedPred && j.doSearch[0])
        {
            WeightGuess g = weightGuess(*fenc, *ref0, param.bitDepth);
            if (!g.skip)
            {
                x265cu_weight_item w0 = { fenc->slot, ref0->slot, 0, 0, 0, 0 };
                x265cu_weight_item w1 = { fenc->slot, ref0->slot, 1, g.curScale, g.mindenom, g.curOffset };
                witems.push_back(w0); witems.push_back(w1);
                wjob.push_back((int)jobs.size());
                guesses.push_back(g);
            }
        }
        int specIdx = -1;
        if (est[i].ahead)
        {
            SpecEstimate e;
            e.fenc = fenc; e.ref0 = ref0; e.ref1 = ref1;
            e.fencNum = fenc->frameNum; e.ref0Num = ref0->frameNum; e.ref1Num = ref1->frameNum;
            e.d0 = d0; e.d1 = d1;
            e.doSearch[0] = j.doSearch[0]; e.doSearch[1] = j.doSearch[1];
            e.usedVersion[0] = usedVersion[0]; e.usedVersion[1] = usedVersion[1];
            e.newVersion[0] = newVersion[0]; e.newVersion[1] = newVersion[1];
            e.costStamp = costStamp;
            e.wref = noWeight; e.wdelta = 0;
            memset(&e.res, 0, sizeof(e.res));
            la.m_spec.push_back(e);
            specIdx = (int)la.m_spec.size() - 1;
            SpecEstimate& se = la.m_spec[specIdx];
            if (!la.m_resident)
            {
                for (int l = 0; l < 2; l++)
                    if (j.doSearch[l])
                    {
                        se.mvs[l].resize((size_t)nCU * 4); se.mvCosts[l].resize((size_t)nCU * 4);
                        j.mvs[l] = &se.mvs[l][0];
                        j.mvCosts[l] = (int32_t*)&se.mvCosts[l][0];
                    }
                se.lowresCosts.resize((size_t)nCU * 2); se.rowSatds.resize((size_t)rows * 4);
                j.lowresCosts = (uint16_t*)&se.lowresCosts[0];
                j.rowSatds = (int32_t*)&se.rowSatds[0];
            }
        }
        else
        {
            for (int l = 0; l < 2; l++)
            {
                int d = l ? d1 : d0;
                if (j.doSearch[l] && !la.m_resident)
                {
                    j.mvs[l] = fenc->lowresMvs[l][d - 1];
                    j.mvCosts[l] = fenc->lowresMvCosts[l][d - 1];
                }
                if (j.doSearch[l]) fenc->mvVersion[l][d - 1] = newVersion[l];
            }
            fenc->costStamp[d0][d1] = costStamp;
            if (!la.m_resident)
            {
                j.lowresCosts = fenc->lowresCosts[d0][d1];
                j.rowSatds = fenc->rowSatds[d0][d1];
            }
        }
        jobs.push_back(j);
        jobOf.push_back(i);
        wref.push_back(noWeight);
        wdelta.push_back(0);
        specOf.push_back(specIdx);
    }
    if (jobs.empty()) return true;
    /* queued cuTree steps read the mirrors as they are now: they go first (same stream, no wait) */
    if (!la.m_ctOps.empty() && !la.cuTreeRun(NULL, 0)) return false;

    /* second half of weightsAnalyse (slicetype.cpp:432-487): the two SATD sweeps run on the GPU */
    if (!witems.empty())
    {
        std::vector<uint32_t> costs(witems.size());
        int r = x265cu_weight_cost_batch(la.m_ctx, (int)witems.size(), &witems[0], &costs[0]);
        if (r) { snprintf(la.m_error, sizeof(la.m_error), "x265cu_weight_cost_batch: %s", x265cu_last_error(la.m_ctx)); return false; }
        for (size_t k = 0; k < wjob.size(); k++)
        {
            x265cu_job& j = jobs[wjob[k]];
            const WeightGuess& g = guesses[k];
            unsigned int origscore = costs[2 * k], minscore = origscore;
            if (!minscore) continue;
            int minscale = g.minscale, mindenom = g.mindenom, minoff = 0, found = 0;
            unsigned int s = costs[2 * k + 1];
            if (s < minscore) { minscore = s; minscale = g.curScale; minoff = g.curOffset; found = 1; }
            while (mindenom > 0 && !(minscale & 1)) { mindenom--; minscale >>= 1; }
            if (!found || (minscale == 1 << mindenom && minoff == 0) || (float)minscore / origscore > 0.998f)
                continue;
            wdelta[wjob[k]] = minscore / origscore;     /* unsigned integer division, as in the reference */
            WeightParam wp = { 1, minscale, mindenom, minoff };
            wref[wjob[k]] = wp;
            j.weighted = 1; j.wScale = minscale; j.wDenom = mindenom; j.wOffset = minoff;
        }
    }

    std::vector<x265cu_job_result> res(jobs.size());
    int r = x265cu_estimate_batch(la.m_ctx, (int)jobs.size(), &jobs[0], &res[0]);
    if (r)
    {
        la.m_spec.resize(specBase);
        snprintf(la.m_error, sizeof(la.m_error), "x265cu_estimate_batch: %s", x265cu_last_error(la.m_ctx));
        return false;
    }
    for (size_t k = 0; k < jobs.size(); k++)
    {
        const int d0 = jobs[k].d0, d1 = jobs[k].d1;
        if (specOf[k] >= 0)
        {
            SpecEstimate& e = la.m_spec[specOf[k]];
            e.res = res[k];
            e.wref = wref[k];
            e.wdelta = wdelta[k];
            la.m_specStats[0]++;
            continue;
        }
        Lowres* fenc = est[jobOf[k]].fenc;
        fenc->weightedRef[d0].present = 0;
        if (wref[k].present) { fenc->weightedRef[d0] = wref[k]; fenc->weightedCostDelta[d0] = wdelta[k]; }
        fenc->costEst[d0][d1] = res[k].costEst;
        fenc->costEstRaw[d0][d1] = res[k].costEstRaw;
        fenc->costEstAq[d0][d1] = res[k].costEstAq;
        if (d1 == 0)
            fenc->intraMbs[d0] += res[k].intraMbs;
        if (la.m_resident)
        {
            /* nothing was copied back: clear the host-side "not searched / not computed" markers */
            if (jobs[k].doSearch[0]) fenc->lowresMvs[0][d0 - 1][0].x = 0;
            if (jobs[k].doSearch[1]) fenc->lowresMvs[1][d1 - 1][0].x = 0;
            fenc->rowSatds[d0][d1][0] = 0;
        }
    }
    return true;
}

/* ================================================================================================
 * cuTree propagation
 * ============================================================================================== */
static inline double clipDuration(double f) { return f < 0.01 ? 0.01 : (f > 1.00 ? 1.00 : f); }   /* CLIP_DURATION, ratecontrol.h:47 */

/* memset(frames[x]->propagateCost, 0, m_cuCount * sizeof(uint16_t)), slicetype.cpp:1668-1701 */
void Lookahead::cuTreeZero(Lowres& f)
{
    x265cu_cutree_op op;
    memset(&op, 0, sizeof(op));
    op.kind = X265CU_CT_ZERO;
    op.fenc = f.slot;
    m_ctOps.push_back(op);
    memset(f.propagateCost, 0, (size_t)m_cuCount * sizeof(uint16_t));
    f.propagateStale = false;       /* host and (once the queue ran) device agree: all zero */
}

/* Lookahead::estimateCUPropagate, slicetype.cpp:1741-1839 (the VBV-only cuTreeFinish at its end is the caller's) */
bool Lookahead::estimateCUPropagate(Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced)
{
    Lowres *fenc = frames[b], *ref0 = frames[p0], *ref1 = frames[p1];
    const int d0 = b - p0, d1 = p1 - b;
    if (!fenc || !ref0 || !ref1 || d0 < 1 || d1 < 0 || d0 > m_param.bframes + 1 || d1 > m_param.bframes + 1)
    {
        snprintf(m_error, sizeof(m_error), "estimateCUPropagate: bad frames (%d, %d, %d)", p0, b, p1);
        return false;
    }
    if (!fenc->invQscaleFactor) { snprintf(m_error, sizeof(m_error), "estimateCUPropagate: cuTree needs the AQ arrays (aq-mode 0)"); return false; }
    /* the kernel reads the device mirrors of lowresCosts[d0][d1] and of the MV fields: they must be the official
     * arrays (an estimate computed ahead and never asked for may have overwritten a mirror) */
    struct { bool stale; int which, a, b; const void* host; } need[3] = {
        { fenc->costStamp[d0][d1] != fenc->devCostStamp[d0][d1], 4, d0, d1, fenc->lowresCosts[d0][d1] },
        { fenc->mvVersion[0][d0 - 1] != fenc->devMvVersion[0][d0 - 1], 6, 0, d0, fenc->lowresMvs[0][d0 - 1] },
        { d1 > 0 && fenc->mvVersion[1][d1 - 1] != fenc->devMvVersion[1][d1 - 1], 6, 1, d1, d1 > 0 ? fenc->lowresMvs[1][d1 - 1] : NULL } };
    for (int i = 0; i < 3; i++)
    {
        if (!need[i].stale) continue;
        /* resident mode has no host copy: the mirror is all there is (the estimate handed out from the look-ahead
         * cache then also only exists as that mirror); counted, reported by the bench, expected to stay 0 */
        if (m_resident) { m_ctStats[2]++; continue; }
        if (!cuTreeRun(NULL, 0)) return false;     /* queued steps read the mirror as it is now */
        if (x265cu_frame_set_array(m_ctx, fenc->slot, need[i].which, need[i].a, need[i].b, need[i].host))
        {
            snprintf(m_error, sizeof(m_error), "x265cu_frame_set_array: %s", x265cu_last_error(m_ctx));
            return false;
        }
        if (i == 0) fenc->devCostStamp[d0][d1] = fenc->costStamp[d0][d1];
        else fenc->devMvVersion[need[i].a][need[i].b - 1] = fenc->mvVersion[need[i].a][need[i].b - 1];
        m_ctStats[2]++;
    }
    x265cu_cutree_op op;
    memset(&op, 0, sizeof(op));
    op.kind = X265CU_CT_PROPAGATE;
    op.fenc = fenc->slot; op.ref0 = ref0->slot; op.ref1 = ref1->slot;
    op.d0 = d0; op.d1 = d1;
    op.referenced = referenced;
    const int distScaleFactor = ((d0 << 8) + ((p1 - p0) >> 1)) / (p1 - p0);
    op.bipredWeight = m_param.bEnableWeightedBiPred ? 64 - (distScaleFactor >> 2) : 32;
    op.fpsFactor = clipDuration((double)m_param.fpsDenom / m_param.fpsNum) / clipDuration(averageDuration);
    m_ctOps.push_back(op);
    ref0->propagateStale = true; m_ctTouched.push_back(ref0);
    if (d1 > 0) { ref1->propagateStale = true; m_ctTouched.push_back(ref1); }
    if (!referenced) { fenc->propagateStale = true; m_ctTouched.push_back(fenc); }    /* its first row is zeroed by the step */
    m_ctStats[0]++;
    return true;
}

/* run the queued steps in one launch.  nFetch != 0: the host wants fetch[0]'s propagateCost now -- every array the
 * pass touched comes back with it (a few 16 KB arrays in the same copy), so that the Lowres arrays are complete
 * after a cuTreeFinish and the pass's second cuTreeFinish (b-pyramid middle frame) needs no launch of its own. */
bool Lookahead::cuTreeRun(Lowres** fetch, int nFetch)
{
    std::vector<int> slots;
    std::vector<uint16_t*> outs;
    std::vector<Lowres*> got;
    if (nFetch)
    {
        for (int i = 0; i < nFetch; i++) m_ctTouched.push_back(fetch[i]);
        for (size_t i = 0; i < m_ctTouched.size(); i++)
        {
            Lowres* l = m_ctTouched[i];
            if (!l->propagateStale) continue;
            l->propagateStale = false;                 /* also de-duplicates */
            slots.push_back(l->slot); outs.push_back(l->propagateCost); got.push_back(l);
        }
        m_ctTouched.clear();
    }
    if (m_ctOps.empty() && slots.empty()) return true;
    int r = x265cu_cutree_run(m_ctx, (int)m_ctOps.size(), m_ctOps.empty() ? NULL : &m_ctOps[0], (int)slots.size(),
                              slots.empty() ? NULL : &slots[0], outs.empty() ? NULL : &outs[0]);
    m_ctOps.clear();
    if (r)
    {
        for (size_t i = 0; i < got.size(); i++) got[i]->propagateStale = true;
        snprintf(m_error, sizeof(m_error), "x265cu_cutree_run: %s", x265cu_last_error(m_ctx));
        return false;
    }
    m_ctStats[1]++;
    return true;
}

const uint16_t* Lookahead::propagateCost(Lowres& f)
{
    Lowres* p = &f;
    if ((f.propagateStale || !m_ctOps.empty()) && !cuTreeRun(&p, 1)) return NULL;
    return f.propagateCost;
}

/* Lookahead::cuTreeFinish, slicetype.cpp:1844-1862: float mapping on the host (compiled like the reference) */
/* the loop of Lookahead::cuTreeFinish (slicetype.cpp:1853-1862).  X265_LOG2 of small positive integers is memoised: the
 * arguments are integers (costs), the same few thousand values come back frame after frame, and a cached result of the same
 * libm call is the same double.  One table per process (zero pages until touched; 0.0 = not computed yet -- log2(1) is simply
 * recomputed); entries are written whole and are the same from every thread. */
void cuTreeFinishMap(const int32_t* intraCost, const int* invQscaleFactor, const uint16_t* propagateCost, const double* qpAqOffset,
                     double* qpCuTreeOffset, int cuCount, int fpsFactor, double weightdelta, double cuTreeStrength)
{
    static uint64_t* const table = (uint64_t*)calloc((size_t)1 << 20, sizeof(uint64_t));
    const int lutSize = table ? 1 << 20 : 0;
    struct Memo
    {
        static inline double log2i(uint64_t* t, int size, int v)
        {
            if (v <= 1 || v >= size) return log2((double)v);
            uint64_t bits = __atomic_load_n(&t[v], __ATOMIC_RELAXED);
            double r;
            if (bits) { memcpy(&r, &bits, sizeof(r)); return r; }
            r = log2((double)v);
            memcpy(&bits, &r, sizeof(r));
            __atomic_store_n(&t[v], bits, __ATOMIC_RELAXED);
            return r;
        }
    };
    for (int cuIndex = 0; cuIndex < cuCount; cuIndex++)
    {
        int intracost = (intraCost[cuIndex] * invQscaleFactor[cuIndex] + 128) >> 8;
        if (intracost)
        {
            int propCost = (propagateCost[cuIndex] * fpsFactor + 128) >> 8;
            const double la = Memo::log2i(table, lutSize, intracost + propCost), lb = Memo::log2i(table, lutSize, intracost);
            volatile double diff = la - lb;          /* evaluated left to right, as the reference's object code does */
            double log2_ratio = diff + weightdelta;
            qpCuTreeOffset[cuIndex] = qpAqOffset[cuIndex] - cuTreeStrength * log2_ratio;
        }
    }
}

bool Lookahead::cuTreeFinish(Lowres* frame, double averageDuration, int ref0Distance)
{
    if (!cuTreeRun(&frame, 1)) return false;
    if (m_resident) return true;         /* intraCost / qpAqOffset are not on the host in resident mode (propagateCost is) */
    int fpsFactor = (int)(clipDuration(averageDuration) / clipDuration((double)m_param.fpsDenom / m_param.fpsNum) * 256);
    double weightdelta = 0.0;
    if (ref0Distance && frame->weightedCostDelta[ref0Distance - 1] > 0)
        weightdelta = (1.0 - frame->weightedCostDelta[ref0Distance - 1]);
    cuTreeFinishMap(frame->intraCost, frame->invQscaleFactor, frame->propagateCost, frame->qpAqOffset, frame->qpCuTreeOffset, m_cuCount,
                    fpsFactor, weightdelta, m_cuTreeStrength);
    return true;
}

} // namespace x265cu


/* ================================================================================================
 * flat C view
 * ============================================================================================== */
using namespace x265cu;

extern "C" {

/* the float mapping of cuTreeFinish on caller-supplied arrays: needs no GPU (CPU test of the pinned-compiler arithmetic) */
void x265cuh_cutree_finish_map(const int32_t* intraCost, const int* invQscaleFactor, const uint16_t* propagateCost, const double* qpAqOffset,
                               double* qpCuTreeOffset, int cuCount, int fpsFactor, double weightdelta, double cuTreeStrength)
{
    cuTreeFinishMap(intraCost, invQscaleFactor, propagateCost, qpAqOffset, qpCuTreeOffset, cuCount, fpsFactor, weightdelta, cuTreeStrength);
}

void* x265cuh_open(const x265cuh_params* p, char* err, int errLen)
{
    Param q;
    memset(&q, 0, sizeof(q));
    q.sourceWidth = p->sourceWidth; q.sourceHeight = p->sourceHeight; q.bitDepth = p->bitDepth; q.maxCUSize = p->maxCUSize;
    q.bframes = p->bframes; q.lookaheadDepth = p->lookaheadDepth; q.lookaheadSlices = p->lookaheadSlices; q.poolWorkers = p->poolWorkers;
    q.fpsNum = p->fpsNum; q.fpsDenom = p->fpsDenom; q.qCompress = p->qCompress; q.bEnableWeightedBiPred = p->bEnableWeightedBiPred;
    q.bBPyramid = 1;         /* x265's default; the flat view is only driven with traces of default-pyramid runs (a hint, see Param) */
    q.bEnableWeightedPred = p->bEnableWeightedPred; q.aqMode = p->aqMode; q.aqStrength = p->aqStrength;
    q.bFrameBias = p->bFrameBias; q.device = p->device; q.frameSlots = p->frameSlots;
    q.stream = p->stream; q.searchWarps = p->searchWarps;
    Lookahead* la = new Lookahead();
    if (!la->create(q))
    {
        if (err && errLen > 0) snprintf(err, errLen, "%s", la->m_error);
        delete la;
        return NULL;
    }
    return la;
}

void x265cuh_close(void* la) { delete (Lookahead*)la; }
int x265cuh_sync(void* la) { return x265cu_sync(((Lookahead*)la)->m_ctx); }
void x265cuh_set_resident(void* la, int on) { ((Lookahead*)la)->m_resident = on != 0; }
int x265cuh_frame_slot(void* frame) { return ((Lowres*)frame)->slot; }
void* x265cuh_ctx(void* la) { return ((Lookahead*)la)->m_ctx; }
const char* x265cuh_error(void* la) { return ((Lookahead*)la)->m_error; }

void x265cuh_info(void* h, int* o)
{
    Lookahead* la = (Lookahead*)h;
    o[0] = la->m_8x8Width; o[1] = la->m_8x8Height; o[2] = la->m_cuCount; o[3] = la->m_geom.stride;
    o[4] = (int)la->m_geom.planeSize; o[5] = la->m_numCoopSlices; o[6] = la->m_numRowsPerSlice; o[7] = la->m_lambda;
    o[8] = la->m_geom.pixelBytes; o[9] = la->m_geom.paddedLines; o[10] = (int)la->m_geom.padOffset; o[11] = la->m_8x8Blocks;
}

uint32_t x265cuh_mvcost_crc(void* h) { return crc32(((Lookahead*)h)->m_mvcost, (4 * 32768 + 1) * sizeof(uint16_t)); }
uint32_t x265cuh_crc32(const void* p, size_t n) { return crc32(p, n); }

void* x265cuh_frame_alloc(void* la) { return ((Lookahead*)la)->allocLowres(); }
void x265cuh_frame_free(void* la, void* f) { ((Lookahead*)la)->freeLowres((Lowres*)f); }

int x265cuh_pre_lookahead(void* la, void* frame, const void* y, intptr_t ys, const void* u, const void* v, intptr_t cs, int poc, int planesBack)
{
    return ((Lookahead*)la)->preLookahead(*(Lowres*)frame, y, ys, u, v, cs, poc, planesBack != 0) ? 0 : -1;
}

int x265cuh_add_pictures(void* la, int n, void** frames, const void* const* y, const intptr_t* ys, const void* const* u, const void* const* v,
                         const intptr_t* cs)
{
    for (int i = 0; i < n; i++)
    {
        Lookahead::PictureIn p = { y[i], ys[i], u[i], v[i], cs[i], 0, NULL };
        if (!((Lookahead*)la)->addPicture(*(Lowres*)frames[i], p)) return -1;
    }
    return 0;
}

int x265cuh_pre_lookahead_batch(void* la, int n, void** frames, const void* const* y, const intptr_t* ys, const void* const* u, const void* const* v,
                                const intptr_t* cs, const int* pocs, int planesBack)
{
    std::vector<Lookahead::PictureIn> pics((size_t)n);
    for (int i = 0; i < n; i++)
    {
        pics[i].y = y[i]; pics[i].yStride = ys[i]; pics[i].u = u[i]; pics[i].v = v[i]; pics[i].cStride = cs[i]; pics[i].poc = pocs[i];
        pics[i].quantOffsets = NULL;
    }
    return ((Lookahead*)la)->preLookaheadBatch(n, (Lowres**)frames, n ? &pics[0] : NULL, planesBack != 0) ? 0 : -1;
}

int x265cuh_estimate(void* h, void** frames, int nframes, const int* triples, int n, int batch, int64_t* scores)
{
    Lookahead* la = (Lookahead*)h;
    (void)nframes;
    CostEstimateGroup grp(*la, (Lowres**)frames);
    if (batch)
    {
        for (int i = 0; i < n; i++)
            grp.add(triples[3 * i], triples[3 * i + 1], triples[3 * i + 2]);
        if (!grp.finishBatch()) return -1;
        for (int i = 0; i < n; i++)
        {
            Lowres* f = (Lowres*)frames[triples[3 * i + 2]];
            scores[i] = f->costEst[triples[3 * i + 2] - triples[3 * i]][triples[3 * i + 1] - triples[3 * i + 2]];
        }
    }
    else
        for (int i = 0; i < n; i++)
        {
            scores[i] = grp.singleCost(triples[3 * i], triples[3 * i + 1], triples[3 * i + 2], false);
            if (scores[i] < 0) return -1;
        }
    return 0;
}

const void* x265cuh_array(void* h, void* frame, int which, int d0, int d1, size_t* bytes)
{
    Lookahead* la = (Lookahead*)h;
    Lowres* l = (Lowres*)frame;
    const size_t n = la->m_cuCount;
    switch (which)
    {
    case 0: *bytes = (size_t)4 * la->m_geom.planeSize * la->m_geom.pixelBytes; return l->buffer[0];
    case 1: *bytes = n * 4; return l->intraCost;
    case 2: *bytes = n; return l->intraMode;
    case 3: *bytes = l->invQscaleFactor ? n * 4 : 0; return l->invQscaleFactor;
    case 4: *bytes = n * 2; return l->lowresCosts[d0][d1];
    case 5: *bytes = (size_t)la->m_8x8Height * 4; return l->rowSatds[d0][d1];
    case 6: *bytes = n * 4; return l->lowresMvs[d0][d1 - 1];
    case 7: *bytes = n * 4; return l->lowresMvCosts[d0][d1 - 1];
    case 8: *bytes = d0 ? (size_t)la->m_8x8Width * 2 : n * 2; return la->propagateCost(*l);
    case 9: *bytes = l->qpCuTreeOffset ? n * 8 : 0; return l->qpCuTreeOffset;
    case 10: *bytes = l->qpAqOffset ? n * 8 : 0; return l->qpAqOffset;
    default: *bytes = 0; return NULL;
    }
}

void x265cuh_cutree_zero(void* la, void* frame) { ((Lookahead*)la)->cuTreeZero(*(Lowres*)frame); }
int x265cuh_cutree_propagate(void* la, void** frames, int nframes, int p0, int p1, int b, int referenced, double averageDuration)
{
    if (p0 < 0 || p1 >= nframes || b < p0 || b > p1) return -1;
    return ((Lookahead*)la)->estimateCUPropagate((Lowres**)frames, averageDuration, p0, p1, b, referenced) ? 0 : -1;
}
int x265cuh_cutree_finish(void* la, void* frame, double averageDuration, int ref0Distance)
{
    return ((Lookahead*)la)->cuTreeFinish((Lowres*)frame, averageDuration, ref0Distance) ? 0 : -1;
}
int x265cuh_cutree_sequence(void* h, int n, const int* kind, void* const* fenc, void* const* ref0, void* const* ref1,
                            const int* d0, const int* d1, const int* arg, const double* averageDuration)
{
    Lookahead* la = (Lookahead*)h;
    Lowres* fr[2 * BFRAME_MAX + 4];
    for (int i = 0; i < n; i++)
    {
        if (kind[i] == 0) la->cuTreeZero(*(Lowres*)fenc[i]);
        else if (kind[i] == 1)
        {
            if (d0[i] < 1 || d1[i] < 0 || d0[i] + d1[i] > 2 * BFRAME_MAX + 2) return -1;
            memset(fr, 0, sizeof(fr));
            fr[0] = (Lowres*)ref0[i]; fr[d0[i] + d1[i]] = (Lowres*)ref1[i]; fr[d0[i]] = (Lowres*)fenc[i];
            if (!la->estimateCUPropagate(fr, averageDuration[i], 0, d0[i] + d1[i], d0[i], arg[i])) return -1;
        }
        else if (!la->cuTreeFinish((Lowres*)fenc[i], averageDuration[i], arg[i])) return -1;
    }
    return 0;
}
void x265cuh_cutree_stats(void* la, int64_t* o) { for (int i = 0; i < 3; i++) o[i] = ((Lookahead*)la)->m_ctStats[i]; }

void x265cuh_frame_scalars(void* frame, int d0, int d1, int64_t* o)
{
    Lowres* l = (Lowres*)frame;
    o[0] = l->costEst[d0][d1]; o[1] = l->costEstAq[d0][d1]; o[2] = l->intraMbs[d0];
    o[3] = (int64_t)l->wp_ssd[0]; o[4] = (int64_t)l->wp_sum[0];
    o[5] = l->weightedRef[d0].present; o[6] = l->weightedRef[d0].scale; o[7] = l->weightedRef[d0].denom;
    o[8] = l->weightedRef[d0].offset;
}

} // extern "C"
