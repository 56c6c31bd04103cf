/* x265_glue.cpp -- the x265-side binding of libx265cu.so, inside the real x265 1.9 encoder.
 *
 * INTEGRATION PROOF (test infrastructure, built by integration/build_x265_cu.py into
 * oracle/_ref/x265_cu<depth>): compiled WITH the reference's headers, linked with the reference's
 * objects and with libx265cu.so.  It performs, on the real `Lowres`/`Lookahead` objects, exactly
 * the edits INTEGRATION.md describes, so that the unchanged host code (slicetypeDecide, scenecut,
 * cuTree, rate control, the whole encoder) consumes what the GPU produced.  The CLI built this way
 * must write the same bitstream as the stock CLI (tests/test_gpu_x265_cli.py).
 */
#include "common.h"
#include "frame.h"
#include "picyuv.h"
#include "lowres.h"
#include "slicetype.h"
#include "bitcost.h"
#include "motion.h"

#include "x265_glue.h"
#include "x265cu.h"

#include <map>
#include <stdio.h>
#include <stdlib.h>
#include <pthread.h>

using namespace X265_NS;

namespace {

struct ExposeBitCost : public BitCost
{
    const uint16_t* table() const { return m_cost; }
};

struct GlueState
{
    x265cu_ctx* ctx;
    std::map<Lowres*, int> slots;
    int nextSlot, numSlots;
};

pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;
std::map<Lookahead*, GlueState> g_states;
__thread int t_weight[4];    /* present, scale, denom, offset: set by weightsAnalyse on this thread */

void die(const char* what, x265cu_ctx* ctx)
{
    fprintf(stderr, "x265 [error]: lookahead GPU path: %s: %s\n", what, x265cu_last_error(ctx));
    abort();      /* no CPU fallback */
}

GlueState& stateOf(Lookahead* la, PicYuv* pic)
{
    std::map<Lookahead*, GlueState>::iterator it = g_states.find(la);
    if (it != g_states.end()) return it->second;
    GlueState st;
    st.nextSlot = 0;
    st.numSlots = la->m_param->lookaheadDepth + la->m_param->bframes + 2 * X265_MAX(la->m_param->frameNumThreads, 1) + 24;
    ExposeBitCost bc;
    bc.setQP(X265_LOOKAHEAD_QP);
    x265cu_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.srcWidth = pic->m_picWidth; cfg.srcHeight = pic->m_picHeight;
    cfg.bitDepth = X265_DEPTH;
    cfg.marginX = pic->m_lumaMarginX; cfg.marginY = pic->m_lumaMarginY;
    cfg.bframes = la->m_param->bframes;
    cfg.numFrameSlots = st.numSlots;
    cfg.numCoopSlices = la->m_numCoopSlices; cfg.numRowsPerSlice = la->m_numRowsPerSlice;
    cfg.bFrameBias = la->m_param->bFrameBias;
    cfg.lookaheadLambda = (int)x265_lambda_tab[X265_LOOKAHEAD_QP];
    cfg.mvcost = bc.table();
    cfg.device = getenv("X265CU_DEVICE") ? atoi(getenv("X265CU_DEVICE")) : 0;
    if (x265cu_open(&cfg, &st.ctx) != X265CU_OK) die("x265cu_open", NULL);
    return g_states[la] = st;
}

int slotOf(GlueState& st, Lowres* l)
{
    std::map<Lowres*, int>::iterator it = st.slots.find(l);
    if (it != st.slots.end()) return it->second;
    if (st.nextSlot >= st.numSlots) { fprintf(stderr, "x265 [error]: lookahead GPU path: out of frame slots\n"); abort(); }
    return st.slots[l] = st.nextSlot++;
}

} // namespace

extern "C" void x265glue_pre(Lookahead* la, Frame* frame)
{
    pthread_mutex_lock(&g_lock);
    GlueState& st = stateOf(la, frame->m_fencPic);
    Lowres& l = frame->m_lowres;
    int slot = slotOf(st, &l);
    pthread_mutex_unlock(&g_lock);
    /* Lowres::init pixel work (lowres.cpp:155-164): planes come back into Lowres::buffer[0] */
    if (x265cu_frame_init(st.ctx, slot, frame->m_fencPic->m_picOrg[0], frame->m_fencPic->m_stride, 0, l.buffer[0])) die("x265cu_frame_init", st.ctx);
    if (x265cu_frame_set_invqscale(st.ctx, slot, l.invQscaleFactor)) die("x265cu_frame_set_invqscale", st.ctx);
    /* lowresIntraEstimate (slicetype.cpp:230-336) */
    x265cu_intra_out o;
    o.intraCost = l.intraCost; o.intraMode = l.intraMode; o.lowresCosts = l.lowresCosts[0][0]; o.rowSatds = l.rowSatds[0][0];
    if (x265cu_intra(st.ctx, slot, &o)) die("x265cu_intra", st.ctx);
    l.costEst[0][0] = o.sums[0];
    l.costEstAq[0][0] = o.sums[1];
    if (x265cu_sync(st.ctx)) die("x265cu_sync", st.ctx);   /* the padded planes have landed */
    t_weight[0] = 0;
}

extern "C" void x265glue_weight(int scale, int denom, int offset)
{
    t_weight[0] = 1; t_weight[1] = scale; t_weight[2] = denom; t_weight[3] = offset;
}

extern "C" int x265glue_estimate(Lookahead* la, Lowres** frames, int p0, int p1, int b, const bool* bDoSearch, int batchMode)
{
    Lowres* fenc = frames[b];
    pthread_mutex_lock(&g_lock);
    GlueState& st = g_states[la];
    x265cu_job j;
    memset(&j, 0, sizeof(j));
    j.fenc = slotOf(st, fenc); j.ref0 = slotOf(st, frames[p0]); j.ref1 = slotOf(st, frames[p1]);
    pthread_mutex_unlock(&g_lock);
    const int d0 = b - p0, d1 = p1 - b;
    j.d0 = d0; j.d1 = d1;
    j.doSearch[0] = bDoSearch[0]; j.doSearch[1] = bDoSearch[1];
    j.sliced = !batchMode;
    if (bDoSearch[0] && fenc->weightedRef[d0].isWeighted)
    {
        /* weightsAnalyse ran on this thread just before (slicetype.cpp:2001-2002) */
        j.weighted = 1; j.wScale = t_weight[1]; j.wDenom = t_weight[2]; j.wOffset = t_weight[3];
    }
    t_weight[0] = 0;
    if (bDoSearch[0]) { j.mvs[0] = fenc->lowresMvs[0][d0 - 1]; j.mvCosts[0] = fenc->lowresMvCosts[0][d0 - 1]; }
    if (bDoSearch[1]) { j.mvs[1] = fenc->lowresMvs[1][d1 - 1]; j.mvCosts[1] = fenc->lowresMvCosts[1][d1 - 1]; }
    j.lowresCosts = fenc->lowresCosts[d0][d1];
    j.rowSatds = fenc->rowSatds[d0][d1];
    x265cu_job_result r;
    if (x265cu_estimate_batch(st.ctx, 1, &j, &r)) die("x265cu_estimate_batch", st.ctx);
    /* the caller scales the B score and stores it (slicetype.cpp:2053-2057) */
    fenc->costEst[d0][d1] = r.costEstRaw;
    fenc->costEstAq[d0][d1] = r.costEstAq;
    if (p1 == b) fenc->intraMbs[d0] += r.intraMbs;
    return 1;
}

extern "C" int x265glue_propagate(Lookahead* la, Lowres** frames, double fpsFactor, int bipredWeight, int p0, int p1, int b, int referenced)
{
    pthread_mutex_lock(&g_lock);
    GlueState& st = g_states[la];
    const int sb = slotOf(st, frames[b]), s0 = slotOf(st, frames[p0]), s1 = slotOf(st, frames[p1]);
    pthread_mutex_unlock(&g_lock);
    /* x265 keeps cuTree's control flow and its memsets, so the arrays it owns are the truth before every step */
    if (referenced && x265cu_frame_set_propagate(st.ctx, sb, frames[b]->propagateCost)) die("x265cu_frame_set_propagate", st.ctx);
    if (x265cu_frame_set_propagate(st.ctx, s0, frames[p0]->propagateCost)) die("x265cu_frame_set_propagate", st.ctx);
    if (p1 != b && x265cu_frame_set_propagate(st.ctx, s1, frames[p1]->propagateCost)) die("x265cu_frame_set_propagate", st.ctx);
    x265cu_cutree_op op;
    memset(&op, 0, sizeof(op));
    op.kind = X265CU_CT_PROPAGATE;
    op.fenc = sb; op.ref0 = s0; op.ref1 = s1;
    op.d0 = b - p0; op.d1 = p1 - b;
    op.referenced = referenced;
    op.bipredWeight = bipredWeight;
    op.fpsFactor = fpsFactor;
    int outSlots[2] = { s0, s1 };
    uint16_t* outs[2] = { frames[p0]->propagateCost, frames[p1]->propagateCost };
    if (x265cu_cutree_run(st.ctx, 1, &op, p1 != b ? 2 : 1, outSlots, outs)) die("x265cu_cutree_run", st.ctx);
    return 1;
}
