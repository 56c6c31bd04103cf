/* x265_glue.h -- call-outs compiled into a temporary GPU-hooked copy of x265 1.9's
 * encoder/slicetype.cpp (see make_gpu_slicetype.py).  INTEGRATION PROOF, test infrastructure:
 * it shows the five edits of INTEGRATION.md working inside the real encoder and lets the
 * encoder-level bit-exactness (slice types, bitstream md5) be checked against the stock binary. */
#ifndef X265_GLUE_H
#define X265_GLUE_H

namespace X265_NS {
class Frame;
class Lookahead;
struct Lowres;
}

extern "C" {
/* after PreLookaheadGroup::processTasks finished a frame on the CPU: redo Lowres::init's pixel work
 * and lowresIntraEstimate on the GPU and OVERWRITE the host arrays with the GPU's results */
void x265glue_pre(X265_NS::Lookahead* la, X265_NS::Frame* frame);
/* weightsAnalyse accepted a weight for (fenc, ref) on this thread */
void x265glue_weight(int scale, int denom, int offset);
/* estimateFrameCost, non-cached branch: run the estimate on the GPU; returns 1 when done */
int x265glue_estimate(X265_NS::Lookahead* la, X265_NS::Lowres** frames, int p0, int p1, int b, const bool* bDoSearch, int batchMode);
/* estimateCUPropagate, instead of its CU loops: one propagate step on the GPU (the cuTree control flow, its memsets
 * and cuTreeFinish stay x265's; the propagateCost arrays of the frames involved travel with the call); returns 1 */
int x265glue_propagate(X265_NS::Lookahead* la, X265_NS::Lowres** frames, double fpsFactor, int bipredWeight, int p0, int p1, int b, int referenced);
}

#endif
