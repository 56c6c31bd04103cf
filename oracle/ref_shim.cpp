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



} // extern "C"
