/* ref_shim.cpp -- thin C shim over the UNMODIFIED x265 1.9 reference (TEST INFRASTRUCTURE ONLY).
 *
 * Built by oracle/build_ref.py into oracle/_ref/libx265ref<depth>.so together with the reference
 * objects.  Nothing under src/ links it.  It exposes the reference's own C primitives
 * (setupCPrimitives table, primitives.cpp:61) so that the oracle restatement can be pinned per
 * primitive, with the recipes of the reference's TestBench (test/pixelharness.cpp,
 * test/intrapredharness.cpp).  The lookahead-only driver and the observation hooks that write the
 * golden traces live in harness/x265_la_driver.cpp (linked into the same library).
 */
#include "common.h"
#include "primitives.h"
#include "param.h"
#include "bitcost.h"
#include "motion.h"
#include "lowres.h"
#include "slice.h"
#include "mv.h"
#include "x265.h"

#include <stdio.h>
#include <string.h>

using namespace X265_NS;

namespace {
struct ExposeBitCost : public BitCost
{
    const uint16_t* table() const { return m_cost; }
};
} // namespace

/* the pixel loops of encoder/weightPrediction.cpp, from the copy with a named namespace (build_ref.py) */
namespace wpref {
struct Cache
{
    const int * intraCost;
    int         numPredDir;
    int         csp;
    int         hshift;
    int         vshift;
    int         lowresWidthInCU;
    int         lowresHeightInCU;
};
void mcLuma(pixel* mcout, Lowres& ref, const MV * mvs);
void mcChroma(pixel* mcout, pixel* src, intptr_t stride, const MV* mvs, const Cache& cache, int height, int width);
uint32_t weightCost(pixel* fenc, pixel* ref, pixel* weightTemp, intptr_t stride, const Cache& cache, int width, int height, WeightParam* w, bool bLuma);
}

/* ================================================================ (1) primitives */
extern "C" {

int x265ref_depth(void) { return X265_DEPTH; }
int x265ref_pixel_bytes(void) { return (int)sizeof(pixel); }

static x265_param* g_setupParam;
void x265ref_setup(void)
{
    if (!g_setupParam)
    {
        g_setupParam = x265_param_alloc();
        x265_param_default(g_setupParam);
        g_setupParam->cpuid = 0;
        g_setupParam->logLevel = X265_LOG_NONE;
        x265_setup_primitives(g_setupParam);
        MotionEstimate::initScales();
    }
}

int x265ref_sad8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return primitives.pu[LUMA_8x8].sad(a, sa, b, sb); }
int x265ref_satd8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return primitives.pu[LUMA_8x8].satd(a, sa, b, sb); }
int x265ref_sa8d8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return primitives.cu[BLOCK_8x8].sa8d(a, sa, b, sb); }
int x265ref_sa8d16x16(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return primitives.cu[BLOCK_16x16].sa8d(a, sa, b, sb); }
void x265ref_sad_x3_8x8(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, intptr_t stride, int32_t* res)
{ primitives.pu[LUMA_8x8].sad_x3(fenc, r0, r1, r2, stride, res); }
void x265ref_sad_x4_8x8(const pixel* fenc, const pixel* r0, const pixel* r1, const pixel* r2, const pixel* r3, intptr_t stride, int32_t* res)
{ primitives.pu[LUMA_8x8].sad_x4(fenc, r0, r1, r2, r3, stride, res); }
void x265ref_pixelavg8x8(pixel* dst, intptr_t ds, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{ primitives.pu[LUMA_8x8].pixelavg_pp(dst, ds, a, sa, b, sb, 32); }
void x265ref_frame_init_lowres(const pixel* src, pixel* d0, pixel* dh, pixel* dv, pixel* dc, intptr_t ss, intptr_t ds, int w, int h)
{ primitives.frameInitLowres(src, d0, dh, dv, dc, ss, ds, w, h); }
void x265ref_extend_border(pixel* pic, intptr_t stride, int w, int h, int mx, int my) { extendPicBorder(pic, stride, w, h, mx, my); }
void x265ref_intra_filter8(const pixel* in, pixel* out) { primitives.cu[BLOCK_8x8].intra_filter(in, out); }
void x265ref_intra_pred8(int mode, pixel* dst, intptr_t ds, const pixel* src, int bFilter)
{ primitives.cu[BLOCK_8x8].intra_pred[mode](dst, ds, src, mode, bFilter); }
void x265ref_weight_pp(const pixel* src, pixel* dst, intptr_t stride, int w, int h, int w0, int round, int shift, int offset)
{ primitives.weight_pp(src, dst, stride, w, h, w0, round, shift, offset); }
uint64_t x265ref_var16(const pixel* p, intptr_t s) { return primitives.cu[BLOCK_16x16].var(p, s); }
uint64_t x265ref_var8(const pixel* p, intptr_t s) { return primitives.cu[BLOCK_8x8].var(p, s); }
/* pu[partitionFromSizes(w, h)].sad / .satd: the 25 luma PU shapes */
int x265ref_pu_sad(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return primitives.pu[partitionFromSizes(w, h)].sad(a, sa, b, sb); }
int x265ref_pu_satd(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb) { return primitives.pu[partitionFromSizes(w, h)].satd(a, sa, b, sb); }
void x265ref_propagate_cost(int* dst, const uint16_t* pin, const int32_t* intra, const uint16_t* inter, const int32_t* invq, const double* fps, int len)
{ primitives.propagateCost(dst, pin, intra, inter, invq, fps, len); }

void x265ref_mvcost_table(uint16_t* out);
int x265ref_lambda_int(void);
int x265ref_lookahead_qp(void) { return X265_LOOKAHEAD_QP; }
int x265ref_lambda_int(void) { return (int)x265_lambda_tab[X265_LOOKAHEAD_QP]; }
/* copies LUT[-65536 .. 65536] (131073 entries) of BitCost::setQP(X265_LOOKAHEAD_QP), bitcost.cpp:31-59 */
void x265ref_mvcost_table(uint16_t* out)
{
    ExposeBitCost bc;
    bc.setQP(X265_LOOKAHEAD_QP);
    memcpy(out, bc.table() - 2 * 32768, (4 * 32768 + 1) * sizeof(uint16_t));
}
int x265ref_exp2fix8(double x) { return x265_exp2fix8(x); }

/* weightPrediction.cpp: mcLuma :59-90, mcChroma :92-166 (4:2:0), weightCost :168-220 */
void x265ref_wp_mc_luma(pixel* const planes[4], intptr_t stride, int width, int lines, const int16_t* mvs, pixel* mcout)
{
    Lowres ref;
    memset(&ref, 0, sizeof(ref));
    for (int i = 0; i < 4; i++) ref.lowresPlane[i] = planes[i];
    ref.isLowres = true;
    ref.lumaStride = stride; ref.width = width; ref.lines = lines;
    wpref::mcLuma(mcout, ref, (const MV*)mvs);
}
void x265ref_wp_mc_chroma(pixel* src, intptr_t stride, const int16_t* mvs, int lowresWidthInCU, int lowresHeightInCU, int height, int width, pixel* mcout)
{
    wpref::Cache c;
    memset(&c, 0, sizeof(c));
    c.csp = X265_CSP_I420; c.hshift = 1; c.vshift = 1;
    c.lowresWidthInCU = lowresWidthInCU; c.lowresHeightInCU = lowresHeightInCU;
    wpref::mcChroma(mcout, src, stride, (const MV*)mvs, c, height, width);
}
uint32_t x265ref_wp_cost(pixel* fenc, pixel* ref, pixel* weightTemp, intptr_t stride, int width, int height, const int32_t* intraCost,
                         int weighted, int scale, int denom, int offset)
{
    wpref::Cache c;
    memset(&c, 0, sizeof(c));
    c.csp = X265_CSP_I420; c.hshift = 1; c.vshift = 1;
    c.intraCost = intraCost;
    WeightParam w;
    w.bPresentFlag = true; w.inputWeight = scale; w.log2WeightDenom = (uint32_t)denom; w.inputOffset = offset;
    return wpref::weightCost(fenc, ref, weightTemp, stride, c, width, height, weighted ? &w : NULL, intraCost != NULL);
}

/* MotionEstimate::motionEstimate on full-resolution planes (encoder/motion.cpp:571-1172, ref->isLowres == false), through the
 * lookahead's luma-only setSourcePU (:165-181).  Items/results have the layout of include/x265cu.h's x265cu_me_item/_result. */
struct RefMeItem { int64_t offset; int16_t mvmin[2], mvmax[2], qmvp[2]; int16_t numCandidates, merange; int16_t mvc[12][2]; };
struct RefMeResult { int16_t mv[2]; int32_t cost; };
namespace { struct MeWithTable : public MotionEstimate { void setTable(uint16_t* centre) { m_cost = centre; } }; }
/* lutCentre != NULL: the search reads this mvcost table (centre pointer) instead of BitCost's own for `qp` -- the search code
 * is untouched, only the protected m_cost pointer is set, so that goldens can be made with a table every machine can rebuild */
void x265ref_motion_estimate_batch(int method, int subme, int qp, int w, int h, pixel* fencPlane, intptr_t fencStride,
                                   pixel* refPlane, intptr_t refStride, int n, const RefMeItem* items, RefMeResult* out, uint16_t* lutCentre)
{
    x265ref_setup();
    MeWithTable me;
    me.init(method, subme, X265_CSP_I400);
    if (lutCentre) me.setTable(lutCentre); else me.setQP(qp);
    ReferencePlanes ref;
    memset(&ref, 0, sizeof(ref));
    ref.fpelPlane[0] = refPlane;
    ref.lumaStride = refStride;
    ref.isLowres = false;
    for (int i = 0; i < n; i++)
    {
        const RefMeItem& it = items[i];
        me.setSourcePU(fencPlane, fencStride, (intptr_t)it.offset, w, h);
        MV mvc[12], outmv;
        for (int k = 0; k < it.numCandidates; k++) mvc[k] = MV(it.mvc[k][0], it.mvc[k][1]);
        out[i].cost = me.motionEstimate(&ref, MV(it.mvmin[0], it.mvmin[1]), MV(it.mvmax[0], it.mvmax[1]), MV(it.qmvp[0], it.qmvp[1]),
                                        it.numCandidates, mvc, it.merange, outmv);
        out[i].mv[0] = outmv.x; out[i].mv[1] = outmv.y;
    }
}
/* LUT[-65536 .. 65536] of BitCost::setQP(qp), and the lambda it was made with */
void x265ref_mvcost_table_qp(int qp, uint16_t* out)
{
    ExposeBitCost bc;
    bc.setQP(qp);
    memcpy(out, bc.table() - 2 * 32768, (4 * 32768 + 1) * sizeof(uint16_t));
}
double x265ref_lambda(int qp) { return x265_lambda_tab[qp]; }


} // extern "C"
