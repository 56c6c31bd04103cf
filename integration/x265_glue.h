/* x265_glue.h -- the x265-side binding of libx265cu.so: call-outs compiled into temporary copies of x265 1.9's
 * encoder/slicetype.cpp, common/lowres.cpp and common/picyuv.cpp (integration/make_gpu_sources.py inserts them at the
 * sites INTEGRATION.md names; no reference source is held here).  With them x265's OWN lookahead -- slicetypeDecide,
 * slicetypeAnalyse, scenecut, slicetypePath, the control flow of cuTree, vbvLookahead, its thread pool -- is the host
 * and every cost estimate, pre-lookahead frame and cuTree propagate step runs on the GPU:
 *
 *   PreLookaheadGroup::processTasks  -> x265glue_pre_list      the whole list in ONE x265cu_pre_lookahead_batch
 *   CostEstimateGroup::finishBatch   -> x265glue_finish_batch  the whole batch in ONE x265cu_estimate_batch
 *   CostEstimateGroup::estimateFrameCost -> x265glue_ensure    singleCost through the look-ahead estimate cache
 *   LookaheadTLD::weightsAnalyse     -> inside the two above   (x265cu_weight_cost_batch; float guesses on the host)
 *   Lookahead::cuTree memsets / estimateCUPropagate / cuTreeFinish -> x265glue_ct_*   queued, one launch per pass
 *   Lowres::create / PicYuv::create  -> x265glue_malloc        pinned + mapped host arrays (results written in place)
 *
 * Built into oracle/_ref/libx265gpu<depth>.so (lookahead-only driver: bench.py e2e, trace parity) and
 * oracle/_ref/x265_cu<depth> (the full x265 CLI: bitstream md5 parity) by integration/build_x265_cu.py. */
#ifndef X265_GLUE_H
#define X265_GLUE_H

#include <stddef.h>

namespace X265_NS {
class Frame;
class Lookahead;
struct Lowres;
}

extern "C" {
/* ---- memory: Lowres::create (lowres.cpp:30-95) and PicYuv::create (picyuv.cpp:51-88) allocate through these */
void*  x265glue_malloc(size_t bytes);
void   x265glue_free(void* p);
/* the allocations of one Lowres::create come out of one pinned arena of this size (x265glue_lowres_bytes) */
void   x265glue_arena_begin(size_t bytes);
void   x265glue_arena_end(void);
size_t x265glue_lowres_bytes(int picWidth, int picHeight, int marginX, int marginY, int bframes, int pixelBytes);
/* 1: the pixel work of Lowres::init (frameInitLowres + 4x extendPicBorder, lowres.cpp:155-164) is the GPU's */
int    x265glue_active(void);

/* ---- Lookahead::create / destroy (slicetype.cpp:586-631) */
void x265glue_open(X265_NS::Lookahead* la);
void x265glue_close(X265_NS::Lookahead* la);
/* ---- PreLookaheadGroup::processTasks (slicetype.cpp:831-856): the list m_preframes[first .. first + n) */
void x265glue_pre_list(X265_NS::Lookahead* la, X265_NS::Frame** frames, int n);
/* ---- CostEstimateGroup::estimateFrameCost (slicetype.cpp:1977-2066), first statement: make the estimate (p0, p1, b)
 * exist in frames[b] (arrays and costEst/costEstAq/intraMbs/weightedCostDelta); the cached branch then returns it */
void x265glue_ensure(X265_NS::Lookahead* la, X265_NS::Lowres** frames, int p0, int p1, int b);
/* ---- CostEstimateGroup::finishBatch (slicetype.cpp:1919-1926): estimates = m_estimates as (p0, b, p1) triples */
int  x265glue_finish_batch(X265_NS::Lookahead* la, X265_NS::Lowres** frames, const int* estimates, int n);
/* ---- Lookahead::cuTree (slicetype.cpp:1640-1739): after each memset of a propagateCost array; around its std::swap */
void x265glue_ct_zero(X265_NS::Lookahead* la, X265_NS::Lowres* frame);
void x265glue_ct_preswap(X265_NS::Lookahead* la, X265_NS::Lowres* a, X265_NS::Lowres* b);
void x265glue_ct_postswap(X265_NS::Lookahead* la, X265_NS::Lowres* a, X265_NS::Lowres* b);
/* ---- Lookahead::estimateCUPropagate (slicetype.cpp:1741-1839) instead of its CU loops; returns 1 */
int  x265glue_propagate(X265_NS::Lookahead* la, X265_NS::Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced);
/* ---- Lookahead::cuTreeFinish (slicetype.cpp:1844-1862), first statement: frame->propagateCost is complete on the host */
void x265glue_ct_fetch(X265_NS::Lookahead* la, X265_NS::Lowres* frame);
/* ... and its last statement (observation only: trace harness) */
void x265glue_ct_finished(X265_NS::Lookahead* la, X265_NS::Lowres* frame, double averageDuration, int ref0Distance);
int  x265glue_ct_finish(X265_NS::Lookahead* la, X265_NS::Lowres* frame, double averageDuration, int ref0Distance);
/* ---- Lookahead::slicetypeDecide (slicetype.cpp:1005), before the mini-GOP is handed to the output queue: the padded
 * lowres planes copied back for weightPrediction.cpp have landed */
void x265glue_sync(X265_NS::Lookahead* la);
/* ---- weightAnalyse (encoder/weightPrediction.cpp:222-505, called by the frame encoders): before the sweep of a plane -- the
 * source frame, its reference, the plane and the lowres vectors (or NULL) the motion-compensated copy is built with; every
 * weightCost of the sweep (:168-220) then is one launch; x265glue_wp_done ends the analysis of the frame (the GPU context
 * serves one weightAnalyse at a time) */
void x265glue_wp_prepare(X265_NS::Frame* frame, X265_NS::Frame* refFrame, int plane, const void* mvs);
int  x265glue_wp_cost(int weighted, int scale, int denom, int offset, unsigned int* cost);
void x265glue_wp_done(void);
/* ---- harness: totals over every context closed so far: contexts, h2d bytes, d2h bytes, kernel launches,
 * look-ahead estimate cache (launched ahead, handed out, computed on demand, requests) */
void x265glue_totals(long long* out8);
}

#endif
