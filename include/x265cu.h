/* x265cu.h -- C ABI of the B200 (sm_100a) lookahead cost-estimation library (libx265cu.so).
 *
 * Drop-in boundary for ONE hot path of x265 1.9: the lookahead cost estimation.  x265 has no
 * plugin API for its lookahead; the seam is the C++ class boundary Lookahead / Lowres /
 * CostEstimateGroup / PreLookaheadGroup (encoder/slicetype.h:98-240) whose data contract is
 * `struct Lowres` (common/lowres.h:107-159).  The host (x265's own slice-type decision, cuTree,
 * rate control) keeps that layer and calls the entry points below instead of its CPU loops; each
 * entry cites the reference site it replaces (paths relative to x265_1.9/source).  INTEGRATION.md
 * shows the binding a maintainer adds on the x265 side.
 *
 * Plain C: pointers, sizes, PODs.  No C++/torch types.  All functions return 0 on success and a
 * negative X265CU_E* code on failure (x265 convention: no exceptions; the caller logs
 * x265cu_last_error() and sets m_aborted).  There is NO CPU fallback: without a CUDA device every
 * call fails with X265CU_ENODEV.
 *
 * Threading: a ctx belongs to one encoder; x265cu_estimate_batch/x265cu_weight_cost_batch are
 * called by one thread at a time (x265 guarantees it via m_sliceTypeBusy, slicetype.cpp:683);
 * the frame_* / intra entries may come from several pool workers (PreLookaheadGroup,
 * slicetype.cpp:831-856) and are serialised inside the library.
 */
#ifndef X265CU_H
#define X265CU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define X265CU_ABI_VERSION 2
#define X265CU_BFRAME_MAX 16

enum
{
    X265CU_OK = 0,
    X265CU_EINVAL = -1,     /* bad argument / unsupported configuration */
    X265CU_ENODEV = -2,     /* no usable CUDA device */
    X265CU_ENOMEM = -3,
    X265CU_ECUDA = -4       /* a CUDA call or kernel failed; see x265cu_last_error */
};

typedef struct x265cu_ctx x265cu_ctx;

/* Replaces the per-encoder state set up by Lookahead::Lookahead / Lookahead::create
 * (encoder/slicetype.cpp:490-591) and the geometry of Lowres::create (common/lowres.cpp:30-48). */
typedef struct x265cu_config
{
    int srcWidth, srcHeight;      /* PicYuv::m_picWidth / m_picHeight (full-resolution luma) */
    int bitDepth;                 /* X265_DEPTH: 8 -> pixel = uint8_t; 10/12 -> pixel = uint16_t */
    int marginX, marginY;         /* PicYuv::m_lumaMarginX / m_lumaMarginY (g_maxCUSize + 32 / + 16) */
    int bframes;                  /* x265_param::bframes (sizes the [bframes+2][bframes+2] tables) */
    int numFrameSlots;            /* number of device mirrors of `Lowres` (lookahead window + a few) */
    int numCoopSlices;            /* Lookahead::m_numCoopSlices   (slicetype.cpp:546-558) */
    int numRowsPerSlice;          /* Lookahead::m_numRowsPerSlice */
    int bFrameBias;               /* x265_param::bFrameBias (B score scale 100/(130+bias), slicetype.cpp:2055) */
    int lookaheadLambda;          /* (int)x265_lambda_tab[X265_LOOKAHEAD_QP] (slicetype.cpp:237) */
    const uint16_t* mvcost;       /* BitCost::s_costs[X265_LOOKAHEAD_QP] centre pointer, valid on
                                     [-65536, 65536] (bitcost.cpp:30-59); copied at open.  Float
                                     math that builds it stays on the host. */
    int device;                   /* CUDA device ordinal */
    void* stream;                 /* cudaStream_t to run on, or NULL: the ctx creates its own */
    int searchWarps;              /* warps per search CTA (0 = default) */
} x265cu_config;

typedef struct x265cu_geometry
{
    int width, lines;             /* Lowres::width / lines (rounded up to multiples of 8) */
    int stride;                   /* Lowres::lumaStride in samples */
    int paddedLines;              /* lines + 2 * marginY */
    int widthInCU, heightInCU, cuCount;
    int64_t planeSize;            /* samples per padded plane; Lowres::buffer[i+1] - buffer[i] */
    int64_t padOffset;            /* Lowres::lowresPlane[i] - buffer[i] */
    int pixelBytes;
} x265cu_geometry;

int  x265cu_abi_version(void);
int  x265cu_device_count(void);
int  x265cu_open(const x265cu_config* cfg, x265cu_ctx** out);
void x265cu_close(x265cu_ctx* ctx);
const char* x265cu_last_error(const x265cu_ctx* ctx);   /* ctx may be NULL: last open() error */
int  x265cu_get_geometry(const x265cu_ctx* ctx, x265cu_geometry* out);
int  x265cu_sync(x265cu_ctx* ctx);
/* pin + map a host array the library will read/write often (Lowres arrays, PicYuv planes); optional, only affects
 * speed: copies become asynchronous, and result arrays of x265cu_estimate_batch whose destination lies in a
 * registered range are written there by the GPU itself instead of being staged and memcpy'd by a host thread */
int  x265cu_host_register(void* ptr, size_t bytes);
int  x265cu_host_unregister(void* ptr);
/* Work buffers a context grew on demand (argument / staging areas of x265cu_estimate_batch and friends) are kept in a
 * process-wide pool when the context closes and handed to the next context of the same device, so that an encoder
 * opened after another one starts warm (cudaMalloc / cudaMallocHost synchronise the device).  This frees the pool. */
void x265cu_trim(void);

/* ---- Lowres::init (common/lowres.cpp:128-165): frame_init_lowres_core (common/pixel.cpp:549-573)
 * + extendPicBorder x4 (pixel.cpp:908-922).  `luma` points at PicYuv::m_picOrg[0]; the caller's
 * plane must be valid for (2*width+1) x (2*lines+1) samples as PicYuv::copyFromPicture guarantees
 * (picyuv.cpp:168-178,287-298).  srcStride in samples.  If planesOut != NULL the four padded
 * planes (4 * planeSize samples, layout of Lowres::buffer[0]) are copied back for the host-side
 * consumers (weightPrediction.cpp); that copy runs behind the compute stream and is
 * complete once x265cu_sync() has returned (none of the lookahead's own decisions read the planes;
 * the encoder reads them after slicetypeDecide).  The per-frame resets of Lowres::init
 * (costEst = -1, MV sentinels ...) stay on the host.  lumaIsDevice: `luma` is a device pointer. */
int x265cu_frame_init(x265cu_ctx* ctx, int slot, const void* luma, intptr_t srcStride, int lumaIsDevice, void* planesOut);

/* invQscaleFactor of the frame (Lowres::invQscaleFactor, produced on the host by
 * calcAdaptiveQuantFrame's float mapping, slicetype.cpp:163-207); NULL = AQ arrays absent. */
int x265cu_frame_set_invqscale(x265cu_ctx* ctx, int slot, const int32_t* invQscale);

/* ---- integer part of LookaheadTLD::acEnergyCu / calcAdaptiveQuantFrame (slicetype.cpp:48-93):
 * pixel_var<16> on luma + pixel_var<8> on Cb/Cr (pixel.cpp:649-666) per 16x16 block of a 4:2:0
 * picture.  energy[block] = summed AC energy of the three planes (u/v may be NULL: luma only),
 * sums[6] = wp_sum[0..2], wp_ssd[0..2] raw accumulations (before the final normalisation at
 * slicetype.cpp:222-227).  Planes are host (or, planesAreDevice != 0, device) pointers padded like
 * PicYuv to a multiple of 16.  Float mapping to QP offsets stays on the host. */
int x265cu_frame_var(x265cu_ctx* ctx, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride,
                     int planesAreDevice, uint32_t* energy, uint64_t sums[6]);

/* ---- x265cu_frame_init + x265cu_frame_var in one call, as PreLookaheadGroup::processTasks runs them
 * back to back (slicetype.cpp:845-849): the luma is uploaded once and both kernels read it. */
int x265cu_frame_init_var(x265cu_ctx* ctx, int slot, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride,
                          int planesAreDevice, void* planesOut, uint32_t* energy, uint64_t sums[6]);

/* ---- the same for a LIST of frames, as PreLookaheadGroup::processTasks receives them (slicetype.cpp:831-856:
 * m_preframes[0..m_jobTotal)): every upload and kernel of the list is enqueued back to back and the host waits
 * once.  Fields as the arguments of x265cu_frame_init_var; sums points at 6 values. */
typedef struct x265cu_frame_in
{
    int slot;
    const void* y; intptr_t yStride;
    const void* u; const void* v; intptr_t cStride;
    int planesAreDevice;
    void* planesOut;
    uint32_t* energy;
    uint64_t* sums;
} x265cu_frame_in;
int x265cu_frame_init_var_batch(x265cu_ctx* ctx, int n, const x265cu_frame_in* items);

/* ---- Lookahead::addPicture (slicetype.cpp:633-650): a picture has arrived in the lookahead's input queue.  Optional.
 * Starts the upload of the picture into a staging area of the slot (asynchronous, on the upload stream) and returns;
 * the pre-lookahead of the frame (x265cu_pre_lookahead_batch / x265cu_frame_init_var_batch with the same pointers and
 * strides for this slot) then finds the picture on the device instead of uploading it while the GPU waits.  The
 * picture must not change between this call and that pre-lookahead; the next initialisation of the slot consumes it. */
int x265cu_frame_upload(x265cu_ctx* ctx, int slot, const void* y, intptr_t yStride, const void* u, const void* v, intptr_t cStride);

/* ---- the whole of PreLookaheadGroup::processTasks for its list (slicetype.cpp:831-856) as ONE pipelined call:
 * x265cu_frame_init_var_batch, the host's float AQ mapping and x265cu_intra_batch, a few frames at a time.  As soon
 * as the energies/sums of frames [first, first + count) are on the host, `aq(user, first, count, invQscale)` is
 * called (on the calling thread, no library call allowed inside): it runs the float part of
 * calcAdaptiveQuantFrame (slicetype.cpp:163-207) for those frames -- on several threads if it likes, as the
 * reference spreads the frames of a list over its workers -- and stores each frame's invQscaleFactor pointer (or
 * NULL: no AQ arrays) in invQscale[0..count); the library publishes them and starts lowresIntraEstimate of those
 * frames on a second stream while the pictures of the later frames are still crossing PCIe.  count is 1 for short
 * lists and up to 8 for long ones.  outs[i] as in x265cu_intra. */
typedef void (*x265cu_aq_fn)(void* user, int first, int count, const int32_t** invQscale);
struct x265cu_intra_out;
int x265cu_pre_lookahead_batch(x265cu_ctx* ctx, int n, const x265cu_frame_in* items, x265cu_aq_fn aq, void* user, struct x265cu_intra_out* outs);

/* ---- LookaheadTLD::lowresIntraEstimate (slicetype.cpp:230-336).  Outputs (any may be NULL):
 * Lowres::intraCost, intraMode, lowresCosts[0][0], rowSatds[0][0]; sums[0] = costEst[0][0],
 * sums[1] = costEstAq[0][0]. */
typedef struct x265cu_intra_out
{
    int32_t* intraCost;
    uint8_t* intraMode;
    uint16_t* lowresCosts;
    int32_t* rowSatds;
    int64_t sums[2];
} x265cu_intra_out;
int x265cu_intra(x265cu_ctx* ctx, int slot, x265cu_intra_out* out);
/* the frames of one PreLookaheadGroup list in one go (one wait) */
int x265cu_intra_batch(x265cu_ctx* ctx, int n, const int* slots, x265cu_intra_out* outs);

/* ---- LookaheadTLD::weightCostLuma (slicetype.cpp:338-371): sum over all 8x8 of
 * min(SATD(weighted ref plane 0, fenc), intraCost).  weighted == 0 measures the plain reference.
 * weightsAnalyse's float guesses and its 0.998 acceptance test stay on the host. */
typedef struct x265cu_weight_item
{
    int fenc, ref;                /* frame slots */
    int weighted;                 /* WeightParam::bPresentFlag */
    int scale, denom, offset;     /* inputWeight, log2WeightDenom, inputOffset */
} x265cu_weight_item;
int x265cu_weight_cost_batch(x265cu_ctx* ctx, int n, const x265cu_weight_item* items, uint32_t* costs);

/* ---- explicit weighted-prediction analysis of the frame encoders, weightAnalyse (encoder/weightPrediction.cpp:222-505):
 * its three pixel loops on planes that are already on the device.  x265cu_wp_prepare names the (slice, list, plane) being
 * analysed: the source frame and its list-0/1 reference by slot, plane 0 = lowres luma (Lowres::lowresPlane, intraCost limits
 * every 8x8 SATD), 1 / 2 = the full-resolution Cb / Cr planes (kept on the device by x265cu_frame_init_var* /
 * x265cu_pre_lookahead_batch).  lowresMvs != NULL (Lowres::lowresMvs[list][diffPoc - 1], host, cuCount vectors): the
 * reference is motion compensated first -- mcLuma (:59-90) / mcChroma (:92-166, 4-tap filters of common/ipfilter.cpp).
 * x265cu_wp_cost then measures weightCost (:168-220) for n candidate weights in one launch (weighted == 0: the plain
 * reference = origscore).  The float guesses, the sweep with its early exits and the acceptance test stay the host's. */
int x265cu_wp_prepare(x265cu_ctx* ctx, int fencSlot, int refSlot, int plane, const void* lowresMvs, const int32_t* intraCost);
int x265cu_wp_cost(x265cu_ctx* ctx, int n, const x265cu_weight_item* cands, uint32_t* costs);

/* ---- CostEstimateGroup::finishBatch / processTasks / estimateFrameCost body
 * (slicetype.cpp:1919-1975, 2004-2051) with estimateCUCost (:2068-2225) and the lowres branch of
 * MotionEstimate::motionEstimate (motion.cpp:571-1172).  One call = one batch (n = 1 for
 * singleCost).  Jobs of one call must be independent, as the reference's batches are. */
typedef struct x265cu_job
{
    int fenc, ref0, ref1;         /* frame slots of frames[b], frames[p0], frames[p1] */
    int d0, d1;                   /* b - p0, p1 - b (d1 == 0: P estimate) */
    int doSearch[2];              /* bDoSearch[] (host derives it from the 0x7FFF sentinel) */
    int sliced;                   /* 1: non-batch estimate -> cooperative slices when the reference
                                     would use them (:2007); 0: whole frame (batch mode) */
    int weighted;                 /* weightedRef[b-p0].isWeighted for the L0 search */
    int wScale, wDenom, wOffset;
    /* host destinations, any may be NULL (result then stays in the device mirror only) */
    void*     mvs[2];             /* Lowres::lowresMvs[l][d-1]   (MV = int16 x, y), written if searched */
    int32_t*  mvCosts[2];         /* Lowres::lowresMvCosts[l][d-1], written if searched */
    uint16_t* lowresCosts;        /* Lowres::lowresCosts[d0][d1] */
    int32_t*  rowSatds;           /* Lowres::rowSatds[d0][d1] */
} x265cu_job;

typedef struct x265cu_job_result
{
    int64_t costEstRaw;           /* sum of bcost over scored CUs */
    int64_t costEst;              /* value the reference stores: raw, or raw*100/(130+bias) for B */
    int64_t costEstAq;
    int32_t intraMbs;             /* to be added to Lowres::intraMbs[d0] when d1 == 0 */
    int32_t reserved;
} x265cu_job_result;
int x265cu_estimate_batch(x265cu_ctx* ctx, int n, const x265cu_job* jobs, x265cu_job_result* results);

/* ---- cuTree propagation: Lookahead::estimateCUPropagate (slicetype.cpp:1741-1839) with its
 * primitive estimateCUPropagateCost (common/pixel.cpp:848-874), and the memsets of Lowres::propagateCost in
 * Lookahead::cuTree (slicetype.cpp:1668-1701).  The device keeps one propagateCost array per frame slot next to
 * the arrays the estimates left there (intraCost, invQscaleFactor, lowresCosts[d0][d1], lowresMvs[l][d]), so a
 * cuTree pass uploads nothing.  One call takes the ops of a (part of a) cuTree pass IN THE REFERENCE'S ORDER and
 * runs them as one launch; outSlots/outPropagateCost name the frames whose propagateCost (cuCount uint16 each)
 * the host wants back (cuTreeFinish reads it, slicetype.cpp:1844-1862; its log2 math stays on the host). */
enum { X265CU_CT_ZERO = 0, X265CU_CT_PROPAGATE = 1 };
typedef struct x265cu_cutree_op
{
    int kind;                     /* X265CU_CT_ZERO: memset(frames[fenc]->propagateCost, 0, cuCount * 2) */
    int fenc, ref0, ref1;         /* frame slots of frames[b], frames[p0], frames[p1] */
    int d0, d1;                   /* b - p0, p1 - b: lowresCosts[d0][d1], lowresMvs[0][d0-1], lowresMvs[1][d1-1] */
    int referenced;               /* estimateCUPropagate's `referenced` */
    int bipredWeight;             /* bipredWeights[0] (slicetype.cpp:1745-1746): 32 without weighted bipred */
    double fpsFactor;             /* CLIP_DURATION(fpsDenom / fpsNum) / CLIP_DURATION(averageDuration) (:1754) */
} x265cu_cutree_op;
int x265cu_cutree_run(x265cu_ctx* ctx, int n, const x265cu_cutree_op* ops, int nOut, const int* outSlots, uint16_t* const* outPropagateCost);
/* overwrite the device's propagateCost of a frame with the host's (hosts that keep the cuTree control flow and
 * its memsets/swaps on their side: INTEGRATION.md) */
int x265cu_frame_set_propagate(x265cu_ctx* ctx, int slot, const uint16_t* propagateCost);
/* overwrite a device mirror with the host's array: which = 4 lowresCosts[d0][d1], 6 lowresMvs[list = d0][d1 - 1]
 * (numbering of the host layer's accessors).  Only needed when the host knows the mirror is not the official array. */
int x265cu_frame_set_array(x265cu_ctx* ctx, int slot, int which, int d0, int d1, const void* data);

/* ---- EncoderPrimitives kernels as batch operations (common/pixel.cpp:39-118,143-322):
 * pu[LUMA_8x8].sad / .satd, cu[BLOCK_8x8].sa8d, cu[BLOCK_16x16].sa8d over n block pairs taken at
 * sample offsets offA[i] / offB[i] of two buffers with strides strideA / strideB (samples).
 * Host arrays in, host array out. */
enum { X265CU_SAD_8x8 = 0, X265CU_SATD_8x8 = 1, X265CU_SA8D_8x8 = 2, X265CU_SA8D_16x16 = 3 };
int x265cu_pixelcmp_batch(x265cu_ctx* ctx, int kind, const void* bufA, size_t samplesA, intptr_t strideA,
                          const void* bufB, size_t samplesB, intptr_t strideB,
                          int n, const int64_t* offA, const int64_t* offB, int32_t* out);
/* pu[partitionFromSizes(width, height)].sad (kind 0) / .satd (kind 1) of the full-resolution motion search (common/pixel.cpp:
 * 954-1004) for any of the 25 luma PU shapes of enum LumaPU (4x4 .. 64x64, the rectangular and the AMP
 * shapes): n block pairs of ONE shape.  sad_x3 / sad_x4 are the same measure with one source offset repeated. */
int x265cu_pixelcmp_pu(x265cu_ctx* ctx, int kind, int width, int height, const void* bufA, size_t samplesA, intptr_t strideA,
                       const void* bufB, size_t samplesB, intptr_t strideB, int n, const int64_t* offA, const int64_t* offB, int32_t* out);
/* MotionEstimate::motionEstimate on full-resolution planes (encoder/motion.cpp:571-1172 with ref->isLowres == false; SURVEY.md
 * 8f-4): n independent searches of ONE PU shape in one launch, a warp per search.  searchMethod = X265_DIA_SEARCH (0),
 * X265_HEX_SEARCH (1), X265_UMH_SEARCH (2), X265_STAR_SEARCH (3), X265_FULL_SEARCH (4) (x265.h); subpelRefine 0..7 selects
 * workload[] (motion.cpp:45-55).  Luma only, as the lookahead's setSourcePU (motion.cpp:165-181; bChromaSATD needs the
 * encoder's Yuv/PicYuv objects and stays out of scope).  An item carries what the reference call takes: setSourcePU's offset
 * (= blockOffset, applied to both planes), mvmin / mvmax in full-pel units, the predictor qmvp and up to 12 candidates mvc in
 * quarter-pel units, merange.  mvcostCentre = BitCost's table of the slice QP (m_cost, bitcost.cpp:31-59; entries
 * [-65536, 65536] are read), built by the host as for x265cu_config::mvcostLUT.  Every block the search can touch -- the
 * window widened by 16 samples on each side -- must lie inside refPlane (x265's planes have the margins for it).
 * Host arrays in, results out; *ms (may be NULL) = the kernel's device time. */
enum { X265CU_DIA_SEARCH = 0, X265CU_HEX_SEARCH = 1, X265CU_UMH_SEARCH = 2, X265CU_STAR_SEARCH = 3, X265CU_FULL_SEARCH = 4 };
typedef struct x265cu_me_item
{
    int64_t offset;
    int16_t mvmin[2], mvmax[2];
    int16_t qmvp[2];
    int16_t numCandidates, merange;
    int16_t mvc[12][2];
} x265cu_me_item;
typedef struct x265cu_me_result { int16_t mv[2]; int32_t cost; } x265cu_me_result;
int x265cu_motion_estimate(x265cu_ctx* ctx, int searchMethod, int subpelRefine, int width, int height,
                           const void* fencPlane, size_t fencSamples, intptr_t fencStride,
                           const void* refPlane, size_t refSamples, intptr_t refStride, const uint16_t* mvcostCentre,
                           int n, const x265cu_me_item* items, x265cu_me_result* out, float* ms);
/* same metric over every aligned 8x8 block of plane 0 of nPairs pairs of frame slots, device
 * resident, ONE launch; out (host, may be NULL) gets nPairs * cuCount results.  Returns the
 * kernel's device time in milliseconds through *ms when ms != NULL (CUDA events on the ctx stream). */
int x265cu_pixelcmp_frames(x265cu_ctx* ctx, int kind, int nPairs, const int* slotsA, const int* slotsB, int32_t* out, float* ms);
/* the same with a plane index per operand (0 = full-pel, 1..3 = the H, V, C half-pel planes lowresMC reads,
 * common/lowres.h:62-81); planesA / planesB may be NULL (plane 0) */
int x265cu_pixelcmp_planes(x265cu_ctx* ctx, int kind, int nPairs, const int* slotsA, const int* planesA, const int* slotsB, const int* planesB,
                           int32_t* out, float* ms);
/* measured integer-issue peaks of this GPU in Gop/s (packed |a-b| accumulate = the SAD inner op;
 * plain 32-bit adds): the roofline the search/cost kernels are judged against */
int x265cu_int_peak(x265cu_ctx* ctx, double* gopsVabsdiff4, double* gopsIadd);

/* ---- instrumentation: device time (ms, CUDA events on the ctx stream) and launch counts
 * accumulated since the last reset, per kernel family. */
enum { X265CU_K_LOWRES = 0, X265CU_K_INTRA, X265CU_K_SEARCH, X265CU_K_COST, X265CU_K_WEIGHT, X265CU_K_PIXEL, X265CU_K_VAR, X265CU_K_CUTREE, X265CU_K_RESULTS, X265CU_K_COUNT };
typedef struct x265cu_stats
{
    double ms[X265CU_K_COUNT];
    int64_t launches[X265CU_K_COUNT];
    int64_t h2dBytes, d2hBytes;
} x265cu_stats;
int x265cu_stats_enable(x265cu_ctx* ctx, int timing);   /* timing != 0: bracket kernels with events */
int x265cu_stats_get(x265cu_ctx* ctx, x265cu_stats* out, int reset);

#ifdef __cplusplus
}
#endif

#endif /* X265CU_H */
