/* x265la_oracle.h -- CPU oracle for the x265 1.9 lookahead cost-estimation path.
 *
 * TEST INFRASTRUCTURE ONLY.  A plain-C restatement of the reference algorithm, written from the
 * behaviour of /root/reference/x265_1.9/source (file:line cited at every function in the .c file).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load it; the product
 * (src/x265_b200) never does.
 *
 * PINNED: every function here is checked (tests/test_oracle_vs_ref.py) against the reference's own
 * compiled C primitives and lookahead (oracle/_ref, built from the unmodified sources by
 * oracle/build_ref.py) and against the golden traces/dumps under tests/golden/ that the reference
 * produced (oracle/gen_golden.py).  The reference ships no golden vectors of its own for this
 * path.
 *
 * Built twice: -DORACLE_DEPTH=8 (pixel = uint8_t) and -DORACLE_DEPTH=10 (pixel = uint16_t).
 */
#ifndef X265LA_ORACLE_H
#define X265LA_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifndef ORACLE_DEPTH
#define ORACLE_DEPTH 8
#endif

#if ORACLE_DEPTH > 8
typedef uint16_t pixel;
#else
typedef uint8_t pixel;
#endif

#define OLA_BFRAME_MAX 16
#define OLA_COST_MAX (1 << 28)
#define OLA_MV_SENTINEL 0x7FFF

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ola_geom
{
    int srcW, srcH;          /* full-resolution luma size */
    int width, lines;        /* lowres size rounded up to a multiple of 8 */
    int stride;              /* lumaStride (samples) */
    int marginX, marginY;
    int paddedLines;         /* lines + 2*marginY */
    int wCU, hCU, nCU;
    int64_t planeSize;       /* samples per padded plane */
    int64_t padOffset;       /* offset of sample (0,0) inside a padded plane */
} ola_geom;

typedef struct ola_mv { int16_t x, y; } ola_mv;

typedef struct ola_frame
{
    ola_geom g;
    int bframes;
    int frameNum;
    int hasAq;
    pixel* buffer[4];
    pixel* plane[4];
    int32_t* intraCost;
    uint8_t* intraMode;
    int32_t* invQscale;
    double* qpAqOffset;
    double* qpCuTreeOffset;
    uint32_t* blockVariance;
    uint16_t* lowresCosts[OLA_BFRAME_MAX + 2][OLA_BFRAME_MAX + 2];
    int32_t* rowSatds[OLA_BFRAME_MAX + 2][OLA_BFRAME_MAX + 2];
    ola_mv* mvs[2][OLA_BFRAME_MAX + 1];
    int32_t* mvCosts[2][OLA_BFRAME_MAX + 1];
    int64_t costEst[OLA_BFRAME_MAX + 2][OLA_BFRAME_MAX + 2];
    int64_t costEstAq[OLA_BFRAME_MAX + 2][OLA_BFRAME_MAX + 2];
    int intraMbs[OLA_BFRAME_MAX + 2];
    uint64_t wp_ssd[3], wp_sum[3];
    uint64_t frameVariance;
    double weightedCostDelta[OLA_BFRAME_MAX + 2];
    uint16_t* propagateCost;    /* Lowres::propagateCost (cuTree) */
} ola_frame;

typedef struct ola_weight
{
    int present;     /* isWeighted */
    int scale;       /* inputWeight */
    int denom;       /* log2WeightDenom */
    int offset;      /* inputOffset (8-bit units) */
} ola_weight;

typedef struct ola_ctx
{
    uint16_t* mvcostBase;    /* LUT[-65536..65536], owned */
    const uint16_t* mvcost;  /* centre pointer */
    int lambda;              /* (int)x265_lambda_tab[X265_LOOKAHEAD_QP] */
    int bFrameBias;
    int numCoopSlices, numRowsPerSlice;
    pixel* wbuffer[4];       /* weighted reference planes (LookaheadTLD::wbuffer) */
    int64_t wplaneSize;
    /* instrumentation (per ola_estimate call) */
    int64_t nSad, nSatd;
} ola_ctx;

/* geometry + lifetime */
int  ola_depth(void);
void ola_geometry(int srcW, int srcH, int marginX, int marginY, ola_geom* g);
ola_frame* ola_frame_create(int srcW, int srcH, int marginX, int marginY, int bframes, int aq);
void ola_frame_destroy(ola_frame* f);
ola_ctx* ola_ctx_create(int bFrameBias, int numCoopSlices, int numRowsPerSlice);
void ola_ctx_destroy(ola_ctx* c);
void ola_mvcost_table(uint16_t* out131073);
int  ola_lambda_int(void);
void ola_coop_slices(int srcH, int lookaheadSlices, int hasPool, int hCU, int* numCoopSlices, int* numRowsPerSlice);

/* primitives */
int  ola_sad8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
int  ola_satd8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
int  ola_satd4x4(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
int  ola_sa8d8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
int  ola_sa8d16x16(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
void ola_pixelavg8x8(pixel* dst, intptr_t ds, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
void ola_frame_init_lowres(const pixel* src, pixel* d0, pixel* dh, pixel* dv, pixel* dc, intptr_t ss, intptr_t ds, int w, int h);
void ola_extend_border(pixel* pic, intptr_t stride, int w, int h, int mx, int my);
void ola_intra_filter8(const pixel* in, pixel* out);
void ola_intra_pred8(int mode, pixel* dst, intptr_t ds, const pixel* src, int bFilter);
void ola_weight_pp(const pixel* src, pixel* dst, intptr_t stride, int w, int h, int w0, int round, int shift, int offset);
uint64_t ola_var16(const pixel* p, intptr_t s);
uint64_t ola_var8(const pixel* p, intptr_t s);
int  ola_exp2fix8(double x);

/* frame-level path */
void ola_frame_init(ola_frame* f, const pixel* srcLuma, intptr_t srcStride, int poc);
void ola_aq_frame(ola_frame* f, const pixel* y, intptr_t ys, const pixel* u, const pixel* v, intptr_t cs,
                  int aqMode, double aqStrength, int weightp);
void ola_intra_estimate(ola_frame* f, int lambda);
void ola_weights_analyse(ola_ctx* c, ola_frame* fenc, ola_frame* ref, ola_weight* out);
uint32_t ola_weight_cost_luma(ola_ctx* c, ola_frame* fenc, ola_frame* ref, const ola_weight* w);
void ola_apply_weight(ola_ctx* c, ola_frame* ref, const ola_weight* w);
/* estimateFrameCost body (non-cached branch); returns the final score stored in costEst.
 * search0/search1 < 0 means "derive from the 0x7FFF sentinel" like the reference.  weight: NULL =
 * run ola_weights_analyse when weightp && search0, else use *weight as given. */
int64_t ola_estimate(ola_ctx* c, ola_frame* fenc, ola_frame* ref0, ola_frame* ref1, int d0, int d1,
                     int search0, int search1, int sliced, int weightp, const ola_weight* weight, ola_weight* usedWeight);

/* cuTree propagation */
void ola_propagate_cost(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                        const int32_t* invQscales, const double* fpsFactor, int len);
void ola_cutree_zero(ola_frame* f);
void ola_estimate_cu_propagate(ola_frame* fenc, ola_frame* ref0, ola_frame* ref1, int d0, int d1, int referenced,
                               double averageDuration, int fpsNum, int fpsDenom, int weightedBiPred);
void ola_cutree_finish(ola_frame* f, double averageDuration, int fpsNum, int fpsDenom, int ref0Distance, double cuTreeStrength);

/* full-resolution PU primitives: pu[LUMA_WxH].sad / .satd for the 25 luma PU
 * shapes (pixel.cpp:39-118, 192-242, 954-1004) */
int ola_pu_sad(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);
int ola_pu_satd(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb);

/* full-resolution motion search of one PU (motion.cpp:571-1172, isLowres == false): searchMethod 0 DIA / 1 HEX / 2 UMH /
 * 3 STAR / 4 FULL, subpelRefine 0..7; vectors: mvmin / mvmax full-pel, qmvp / mvc / result quarter-pel.  mvcostCentre is the
 * centre of a BitCost table of any QP: an input (the host builds it, as for the lookahead) */
typedef struct ola_me_item
{
    int64_t offset;               /* setSourcePU's offset = blockOffset, in samples, into both planes */
    int16_t mvmin[2], mvmax[2];
    int16_t qmvp[2];
    int16_t numCandidates, merange;
    int16_t mvc[12][2];
} ola_me_item;
typedef struct ola_me_result { int16_t mv[2]; int32_t cost; } ola_me_result;
int ola_motion_estimate_pu(int searchMethod, int subpelRefine, int w, int h, const pixel* fencPlane, intptr_t fencStride,
                           const pixel* refPlane, intptr_t refStride, intptr_t offset, const uint16_t* mvcostCentre,
                           const int16_t mvmin[2], const int16_t mvmax[2], const int16_t qmvp[2],
                           int numCandidates, const int16_t* mvc, int merange, int16_t outQMv[2]);
void ola_motion_estimate_batch(int searchMethod, int subpelRefine, int w, int h, const pixel* fencPlane, intptr_t fencStride,
                               const pixel* refPlane, intptr_t refStride, const uint16_t* mvcostCentre, int n, const ola_me_item* items, ola_me_result* out);

/* helpers for tests */
void ola_lowres_mc(pixel* const planes[4], intptr_t stride, intptr_t blockOffset, int qx, int qy, pixel* blk);
/* explicit weighted-prediction analysis, pixel loops (encoder/weightPrediction.cpp:59-220) */
void ola_wp_mc_luma(pixel* const planes[4], intptr_t stride, int width, int lines, const int16_t* mvs, pixel* mcout);
void ola_wp_mc_chroma(const pixel* src, intptr_t stride, const int16_t* mvs, int lowresWidthInCU, int lowresHeightInCU,
                      int height, int width, pixel* mcout);
uint32_t ola_wp_cost(const pixel* fenc, const pixel* ref, pixel* weightTemp, intptr_t stride, int width, int height,
                     const int32_t* intraCost, int weighted, int scale, int denom, int offset);
uint32_t ola_crc32(const void* p, size_t n);
void ola_synth_frame(int w, int h, int t, int nframes, uint32_t seed, void* y, int ystride, void* u, void* v, int cstride);
void ola_copy_picture(const pixel* src, int w, int h, pixel* dst, intptr_t dstStride);

#ifdef __cplusplus
}
#endif

#endif
