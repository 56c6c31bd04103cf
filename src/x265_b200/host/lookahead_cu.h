/* lookahead_cu.h -- host-side lookahead layer above the C ABI (libx265cu.so), in x265's own
 * language (C++), mirroring the reference's interface for this path so that callers and tests
 * read like x265 (names, argument meaning and error behaviour):
 *
 *   x265cu::Lowres              <-> struct Lowres                  common/lowres.h:107-159
 *   x265cu::Lookahead           <-> class Lookahead (cost side)    encoder/slicetype.h:98-176
 *   ::preLookahead()            <-> PreLookaheadGroup::processTasks encoder/slicetype.cpp:831-856
 *   ::calcAdaptiveQuantFrame()  <-> LookaheadTLD::calcAdaptiveQuantFrame  slicetype.cpp:95-228
 *   ::weightsAnalyse()          <-> LookaheadTLD::weightsAnalyse   slicetype.cpp:391-488
 *   x265cu::CostEstimateGroup   <-> class CostEstimateGroup        encoder/slicetype.h:199-240
 *   ::cuTreeZero/estimateCUPropagate/cuTreeFinish <-> the memsets of Lookahead::cuTree (slicetype.cpp:1668-1701),
 *                                   Lookahead::estimateCUPropagate (:1741-1839), Lookahead::cuTreeFinish (:1844-1862)
 *
 * Everything integer and data-parallel runs on the GPU through the C ABI; the small float decisions
 * (AQ strength mapping, weight guesses, mvcost table) stay here on the host, compiled with the
 * reference's flags (-ffast-math) because the bitstream depends on them.
 * The slice-type decision, scenecut, the control flow of cuTree and the VBV logic of x265 are consumers of this
 * layer and are NOT re-implemented (out of scope: they stay x265's own code); of cuTree, the
 * per-CU propagation runs on the GPU, its log2 mapping to QP offsets stays on the host.
 */
#ifndef X265CU_LOOKAHEAD_CU_H
#define X265CU_LOOKAHEAD_CU_H

#include <stdint.h>
#include <stddef.h>
#include <map>
#include <vector>

#include "../../../include/x265cu.h"

namespace x265cu {

enum { BFRAME_MAX = 16, MAX_BATCH_SIZE = 512 };

struct MV { int16_t x, y; };

/* the subset of x265_param the lookahead cost path reads */
struct Param
{
    int sourceWidth, sourceHeight;
    int bitDepth;            /* X265_DEPTH of the build being replaced */
    int maxCUSize;           /* g_maxCUSize (margins) */
    int bframes;
    int lookaheadDepth;
    int lookaheadSlices;     /* as given by the user; normalised like Lookahead::Lookahead */
    int poolWorkers;         /* size of the worker pool the host encoder would have had (slices need a pool) */
    int bEnableWeightedPred;
    int aqMode;              /* 0 none, 1 variance, 2 auto-variance, 3 auto-variance biased */
    double aqStrength;
    int bFrameBias;
    int device;
    int frameSlots;          /* 0 = lookaheadDepth + bframes + 8 */
    void* stream;            /* cudaStream_t to run on (NULL: own stream) */
    int searchWarps;         /* 0 = default */
    /* cuTree (0 = the x265 defaults: 30/1 fps, qcomp 0.6) */
    int fpsNum, fpsDenom;
    double qCompress;        /* rc.qCompress: m_cuTreeStrength = 5 * (1 - qCompress), slicetype.cpp:511 */
    int bEnableWeightedBiPred;
    /* cooperative-slice geometry as the host computed it (Lookahead::Lookahead, slicetype.cpp:534-558); 0 = derive it here
     * from lookaheadSlices / poolWorkers / sourceHeight */
    int forceCoopSlices, forceRowsPerSlice;
    int bBPyramid;           /* x265_param::bBPyramid: only a hint for which estimates to compute ahead, never changes a result */
};

struct WeightParam { int present, scale, denom, offset; };

/* lookahead outputs of one frame: same members, meaning and layout as the reference's Lowres */
struct Lowres
{
    int frameNum;
    int slot;                /* device mirror index */
    int width, lines;
    intptr_t lumaStride;
    int bframes;
    void* buffer[4];         /* padded planes (pixel = uint8_t / uint16_t) */
    void* lowresPlane[4];
    int64_t costEst[BFRAME_MAX + 2][BFRAME_MAX + 2];
    int64_t costEstAq[BFRAME_MAX + 2][BFRAME_MAX + 2];
    int64_t costEstRaw[BFRAME_MAX + 2][BFRAME_MAX + 2];   /* before the B-frame scaling (slicetype.cpp:2053-2057); observation only */
    int32_t* rowSatds[BFRAME_MAX + 2][BFRAME_MAX + 2];
    int intraMbs[BFRAME_MAX + 2];
    int32_t* intraCost;
    uint8_t* intraMode;
    uint16_t* lowresCosts[BFRAME_MAX + 2][BFRAME_MAX + 2];
    int32_t* lowresMvCosts[2][BFRAME_MAX + 1];
    MV* lowresMvs[2][BFRAME_MAX + 1];
    double* qpAqOffset;
    double* qpCuTreeOffset;
    int* invQscaleFactor;
    uint32_t* blockVariance;
    uint64_t wp_ssd[3];
    uint64_t wp_sum[3];
    uint64_t frameVariance;
    double weightedCostDelta[BFRAME_MAX + 2];
    WeightParam weightedRef[BFRAME_MAX + 2];   /* reference keeps ReferencePlanes here; we keep the weight */
    uint16_t* propagateCost; /* cuTree; on the host only as far as it was fetched (Lookahead::propagateCost()) */
    uint8_t* arena;          /* one pinned allocation holding every array above */
    size_t arenaBytes;
    /* bookkeeping of the look-ahead estimate cache (Lookahead::m_spec): the frame went through preLookahead();
     * which search produced lowresMvs[l][d] (0: none yet) */
    bool ready;
    uint64_t mvVersion[2][BFRAME_MAX + 1];
    /* which computation last wrote the DEVICE mirror of lowresMvs[l][d] / lowresCosts[d0][d1], and which one the
     * official (host) lowresCosts[d0][d1] came from: cuTree reads the mirrors and must read the official arrays */
    uint64_t devMvVersion[2][BFRAME_MAX + 1];
    uint64_t costStamp[BFRAME_MAX + 2][BFRAME_MAX + 2], devCostStamp[BFRAME_MAX + 2][BFRAME_MAX + 2];
    bool propagateStale;     /* the device holds a newer propagateCost than the host array */
};

/* One estimate computed AHEAD of its request (see CostEstimateGroup::singleCost).  The reference asks for
 * the non-batch estimates of a slicetypeDecide call one by one, each a GPU round trip with a mostly idle
 * GPU; their order repeats from mini-GOP to mini-GOP, so the estimates that are about to be asked for are
 * launched together with the first one.  An entry is only ever handed out when the request that arrives
 * is EXACTLY the computation that was done (same frames, same bDoSearch[] state, same MV fields consumed):
 * a function of identical inputs, hence bit-identical to computing it on request. */
struct SpecEstimate
{
    Lowres *fenc, *ref0, *ref1;
    int fencNum, ref0Num, ref1Num;       /* frameNum of the three at compute time (slots are recycled) */
    int d0, d1;
    int doSearch[2];
    uint64_t usedVersion[2];             /* version of lowresMvs[l][d] consumed (lists not searched by this estimate) */
    uint64_t newVersion[2];              /* version of the fields this estimate's searches produced */
    uint64_t costStamp;                  /* id of this computation (Lowres::costStamp) */
    x265cu_job_result res;
    WeightParam wref;
    double wdelta;
    std::vector<uint8_t> mvs[2], mvCosts[2], lowresCosts, rowSatds;   /* host copies (not in resident mode) */
};

class Lookahead
{
public:
    Param m_param;
    x265cu_ctx* m_ctx;
    x265cu_geometry m_geom;
    int m_8x8Width, m_8x8Height, m_8x8Blocks, m_cuCount;
    int m_numCoopSlices, m_numRowsPerSlice;
    bool m_bAdaptiveQuant;
    uint16_t* m_mvcost;      /* BitCost::s_costs[X265_LOOKAHEAD_QP] (base pointer) */
    int m_lambda;
    char m_error[512];
    std::vector<int> m_freeSlots;   /* device mirror slots not bound to a Lowres */
    bool m_resident;         /* inputs are device pointers and result arrays stay in HBM (only sums return) */
    /* look-ahead estimate cache */
    bool m_lookAhead;                          /* on by default; X265CU_LOOKAHEAD_CACHE=0 turns it off (experiments) */
    std::map<int, Lowres*> m_byPoc;            /* frames that went through preLookahead(), by frameNum */
    std::vector<SpecEstimate> m_spec;
    struct Request { int p0, b, p1; };         /* frameNums */
    std::vector<Request> m_episode;            /* non-batch estimates asked for since the last batch ... */
    std::vector<std::vector<Request> > m_history;   /* ... and the runs before it (most recent last) */
    std::vector<int> m_historyNewest;          /* how many frames had gone through preLookahead() when each of those runs began */
    int m_newestReady, m_episodeNewest;        /* ... now / when the current run began (a stream's frames arrive in order) */
    bool m_trellisAtBatch, m_trellisBref;      /* X265CU_TRELLIS_AHEAD bits 1, 2 (experiments, off: measured slower): the rule also at batch time / with slicetypeDecide's B-ref variant */
    bool m_trellisAhead;                       /* X265CU_TRELLIS_AHEAD=0 turns the rule-based part of the cache off (experiments) */
    int m_batchFirst, m_batchLast;             /* frameNum range of the most recent batch (-1: none yet) */
    uint64_t m_versionCounter;
    int64_t m_specStats[4];                    /* launched ahead, handed out, requests computed alone, requests total */

    Lookahead();
    ~Lookahead();
    bool create(const Param& p);
    void destroy();

    Lowres* allocLowres();                 /* Lowres::create */
    /* a Lowres whose arrays are the CALLER's (x265's own `struct Lowres`): every pointer member of `arrays`
     * is taken as is, nothing is allocated or registered here (the caller pins its arrays if it wants asynchronous,
     * in-place results); scalars (costEst ...) live in the returned object and are the caller's to mirror */
    Lowres* adoptLowres(const Lowres& arrays);
    void freeLowres(Lowres* l);
    void forgetFrame(Lowres* l);            /* drop everything the look-ahead estimate cache derived from this frame */
    /* Lowres::init (lowres.cpp:128-165); luma = PicYuv::m_picOrg[0] padded as copyFromPicture does */
    bool lowresInit(Lowres& l, const void* luma, intptr_t stride, int poc, bool copyPlanesBack);
    void aqMapRows(Lowres& l, const uint32_t* energy, const float* quantOffsets, int byFirst, int byLast);
    void aqFrameSums(Lowres& l, const uint64_t* sums);
    /* LookaheadTLD::calcAdaptiveQuantFrame; planes padded like PicYuv */
    bool calcAdaptiveQuantFrame(Lowres& l, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride,
                                const uint32_t* preEnergy = NULL, const uint64_t* preSums = NULL, bool publish = true,
                                const float* quantOffsets = NULL);
    void lowresReset(Lowres& l, int poc);   /* the host-side resets of Lowres::init */
    /* the padded planes of preLookahead(copyPlanesBack) are complete on the host after sync() */
    bool sync() { return x265cu_sync(m_ctx) == 0; }
    bool lowresIntraEstimate(Lowres& l);
    /* PreLookaheadGroup::processTasks for one frame */
    bool preLookahead(Lowres& l, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride, int poc, bool copyPlanesBack,
                      const float* quantOffsets = NULL);

    /* PreLookaheadGroup::processTasks for a list of frames: one wait per stage instead of two per frame */
    struct PictureIn { const void* y; intptr_t yStride; const void* u; const void* v; intptr_t cStride; int poc;
                       const float* quantOffsets;   /* Frame::m_quantOffsets (x265_picture::quantOffsets), or NULL */ };
    bool preLookaheadBatch(int n, Lowres** frames, const PictureIn* pics, bool copyPlanesBack);
    /* Lookahead::addPicture (slicetype.cpp:633-650): the picture has arrived in the input queue; its upload starts now
     * (asynchronous) instead of when its pre-lookahead runs.  Optional; the picture must not change until then. */
    bool addPicture(Lowres& l, const PictureIn& pic);

    /* ---- cuTree propagation.  The three calls replace, one to one, what Lookahead::cuTree does
     * to Lowres::propagateCost: its memsets (slicetype.cpp:1668-1701), estimateCUPropagate (:1741-1839) and
     * cuTreeFinish (:1844-1862).  Zero and propagate steps are only QUEUED (they run on the device in one launch
     * per run of steps, in order); cuTreeFinish and propagateCost() run the queue and fetch what they read. */
    std::vector<x265cu_cutree_op> m_ctOps;
    std::vector<Lowres*> m_ctTouched;
    double m_cuTreeStrength;
    int64_t m_ctStats[3];                   /* propagate steps, launches (queue runs), mirrors re-uploaded */
    void cuTreeZero(Lowres& f);
    bool estimateCUPropagate(Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced);
    bool cuTreeFinish(Lowres* frame, double averageDuration, int ref0Distance);
    bool cuTreeRun(Lowres** fetch, int nFetch);
    const uint16_t* propagateCost(Lowres& f);   /* the host copy, fetched if the device is ahead */

    int64_t ncu() const { return m_8x8Blocks; }
    static void mvcostTable(int bitDepth, uint16_t* out131073, int* lambdaInt);
};

/* same call protocol as the reference's CostEstimateGroup: add()/finishBatch() for batches,
 * singleCost() for one estimate (cached results are returned without touching the GPU) */
class CostEstimateGroup
{
public:
    Lookahead& m_lookahead;
    Lowres** m_frames;
    bool m_batchMode;
    struct Estimate { int p0, b, p1; } m_estimates[MAX_BATCH_SIZE];
    int m_jobTotal;

    CostEstimateGroup(Lookahead& l, Lowres** f) : m_lookahead(l), m_frames(f), m_batchMode(false), m_jobTotal(0) {}
    void add(int p0, int p1, int b);
    bool finishBatch();
    int64_t singleCost(int p0, int p1, int b, bool intraPenalty = false);

protected:
    struct EstReq { Lowres *fenc, *ref0, *ref1; int d0, d1; bool ahead; bool sliced; };
    bool runEstimates(const EstReq* e, int n);
    void predictFrom(const std::vector<Lookahead::Request>& ep, size_t first, int shift, Lowres* skipFenc, int skipD0, int skipD1,
                     const std::vector<EstReq>& already, std::vector<EstReq>& out, bool& valid);
    void predictTrellis(const Lookahead::Request& rq, Lowres* skipFenc, int skipD0, int skipD1, const std::vector<EstReq>& already, std::vector<EstReq>& out);
    bool takeAhead(Lowres* fenc, Lowres* ref0, Lowres* ref1, int d0, int d1);
};

} // namespace x265cu

/* ---- flat C view of the layer for ctypes (tests, bench.py): handles are opaque pointers ---- */
extern "C" {
typedef struct x265cuh_params
{
    int sourceWidth, sourceHeight, bitDepth, maxCUSize, bframes, lookaheadDepth, lookaheadSlices, poolWorkers;
    int bEnableWeightedPred, aqMode;
    double aqStrength;
    int bFrameBias, device, frameSlots;
    void* stream;
    int searchWarps;
    int fpsNum, fpsDenom;
    double qCompress;
    int bEnableWeightedBiPred;
} x265cuh_params;
void* x265cuh_open(const x265cuh_params* p, char* err, int errLen);
void  x265cuh_cutree_finish_map(const int32_t* intraCost, const int* invQscaleFactor, const uint16_t* propagateCost, const double* qpAqOffset,
                                double* qpCuTreeOffset, int cuCount, int fpsFactor, double weightdelta, double cuTreeStrength);
void  x265cuh_close(void* la);
int   x265cuh_sync(void* la);                                 /* x265cu_sync: pending plane copy-backs have landed */
void  x265cuh_set_resident(void* la, int on);                /* device-resident inputs/outputs (see Lookahead::m_resident) */
int   x265cuh_frame_slot(void* frame);
void* x265cuh_ctx(void* la);                                  /* the underlying x265cu_ctx* */
void  x265cuh_info(void* la, int* out16);                     /* wCU, hCU, nCU, stride, planeSize(lo32), numCoopSlices, numRowsPerSlice, lambda, pixelBytes */
uint32_t x265cuh_mvcost_crc(void* la);
void* x265cuh_frame_alloc(void* la);
void  x265cuh_frame_free(void* la, void* frame);
int   x265cuh_pre_lookahead(void* la, void* frame, const void* y, intptr_t ys, const void* u, const void* v, intptr_t cs, int poc, int planesBack);
int   x265cuh_pre_lookahead_batch(void* la, int n, void** frames, const void* const* y, const intptr_t* ys, const void* const* u, const void* const* v,
                                  const intptr_t* cs, const int* pocs, int planesBack);
int   x265cuh_add_pictures(void* la, int n, void** frames, const void* const* y, const intptr_t* ys, const void* const* u, const void* const* v,
                           const intptr_t* cs);
/* jobs: n triples (p0, p1, b) as indices into frames[]; batch != 0 -> add()+finishBatch(), else singleCost() each */
int   x265cuh_estimate(void* la, void** frames, int nframes, const int* triples, int n, int batch, int64_t* scores);
/* cuTree: the three calls of x265cu::Lookahead (frames[] indexed like the reference's: p0, p1, b) */
void  x265cuh_cutree_zero(void* la, void* frame);
int   x265cuh_cutree_propagate(void* la, void** frames, int nframes, int p0, int p1, int b, int referenced, double averageDuration);
int   x265cuh_cutree_finish(void* la, void* frame, double averageDuration, int ref0Distance);
void  x265cuh_cutree_stats(void* la, int64_t* out3);
/* a run of those calls in one go (pre-marshalled replays: the per-call cost of the caller's language stays out of the
 * measurement): kind 0 = zero(fenc), 1 = propagate(ref0, fenc, ref1 at distances d0, d1; arg = referenced),
 * 2 = finish(fenc; arg = ref0Distance) */
int   x265cuh_cutree_sequence(void* la, int n, const int* kind, void* const* fenc, void* const* ref0, void* const* ref1,
                              const int* d0, const int* d1, const int* arg, const double* averageDuration);
/* array accessors for checks: which = 0 planes, 1 intraCost, 2 intraMode, 3 invQscale, 4 lowresCosts[d0][d1],
 * 5 rowSatds[d0][d1], 6 lowresMvs[list=d0][d1-1], 7 lowresMvCosts[list=d0][d1-1], 8 propagateCost (fetched from the
 * device if it is ahead; d0 != 0: only the first CU row), 9 qpCuTreeOffset, 10 qpAqOffset; returns pointer and byte size */
const void* x265cuh_array(void* la, void* frame, int which, int d0, int d1, size_t* bytes);
void  x265cuh_frame_scalars(void* frame, int d0, int d1, int64_t* out9);  /* costEst, costEstAq, intraMbs[d0], wp_ssd0, wp_sum0, weighted, scale, denom, offset */
uint32_t x265cuh_crc32(const void* p, size_t n);
const char* x265cuh_error(void* la);
}

#endif
