/* ref_hooks.h -- observation call-outs compiled into the hooked temporary copy of the
 * reference's slicetype.cpp (see make_hooked_slicetype.py).  TEST INFRASTRUCTURE ONLY.
 * Implemented in oracle/ref_shim.cpp.  They record WHICH estimates the reference lookahead ran
 * and checksums of what it produced; they never modify reference state. */
#ifndef X265LA_REF_HOOKS_H
#define X265LA_REF_HOOKS_H

namespace X265_NS {
class Frame;
struct Lowres;
}

extern "C" {
/* slicetype.cpp:851 -- a frame finished Lowres::init + AQ + lowresIntraEstimate */
void x265ref_hook_pre(X265_NS::Frame* frame);
/* slicetype.cpp:1919/1925 -- CostEstimateGroup::finishBatch entry (begin=1) / exit (begin=0) */
void x265ref_hook_batch(int begin, int njobs);
/* slicetype.cpp:2053 -- estimateFrameCost computed a non-cached estimate */
void x265ref_hook_job(X265_NS::Lowres** frames, int p0, int p1, int b, int search0, int search1, int batchMode, int sliced);
/* slicetype.cpp:486 -- weightsAnalyse accepted a weight for (fenc, ref) */
void x265ref_hook_weight(int fencPoc, int refPoc, int scale, int denom, int offset);
/* slicetype.cpp:1668-1701 -- Lookahead::cuTree zeroed a frame's propagateCost */
void x265ref_hook_ctzero(X265_NS::Lowres* frame);
/* slicetype.cpp:1839 -- Lookahead::estimateCUPropagate finished its scatter loop */
void x265ref_hook_propagate(X265_NS::Lowres** frames, double averageDuration, int p0, int p1, int b, int referenced);
/* slicetype.cpp:1862 -- Lookahead::cuTreeFinish wrote qpCuTreeOffset */
void x265ref_hook_ctfinish(X265_NS::Lowres* frame, double averageDuration, int ref0Distance);
}

#endif
