#!/usr/bin/env python3
"""Make temporary GPU-bound copies of three x265 1.9 sources (the binding of INTEGRATION.md, applied by anchors).

usage: make_gpu_sources.py <reference source dir> <out dir>
  encoder/slicetype.cpp -> <out>/slicetype_gpu.cpp
  common/lowres.cpp     -> <out>/lowres_gpu.cpp
  common/picyuv.cpp     -> <out>/picyuv_gpu.cpp
  encoder/weightPrediction.cpp -> <out>/weightpred_gpu.cpp

Holds no reference source: it finds one-line anchors in the files it is given and inserts call-outs to
integration/x265_glue.h.  The copies live under oracle/_ref/ (git-ignored).  What changes in x265 (reference line numbers
of x265_1.9/source):

  slicetype.cpp
    :590  Lookahead::create            + x265glue_open(this)                 one x265cu context per encoder
    :618  Lookahead::destroy           + x265glue_close(this)
    :837  PreLookaheadGroup::processTasks  the first thread to arrive takes the WHOLE list: x265glue_pre_list (Lowres::init
                                       resets, then one x265cu_pre_lookahead_batch: lowres planes, AQ variance, intra)
    :1005 slicetypeDecide              + x265glue_sync(this) before the mini-GOP goes to the output queue
    :1668-1701, :1672, :1734  cuTree   + x265glue_ct_zero after each propagateCost memset; pre/post-swap around std::swap
    :1760 estimateCUPropagate          CU loops replaced by x265glue_propagate (queued)
    :1845 cuTreeFinish                 + x265glue_ct_finish first (runs the queued pass, one launch; log2 mapping in the host layer, then returns), x265glue_ct_finished last
    :1921 CostEstimateGroup::finishBatch   body replaced by x265glue_finish_batch (one x265cu_estimate_batch per batch)
    :1980 estimateFrameCost            + x265glue_ensure first (singleCost: look-ahead estimate cache, weightsAnalyse on the GPU)
  lowres.cpp
    :32   Lowres::create               arrays from one pinned arena (x265glue_arena_begin/end, x265_malloc -> x265glue_malloc)
    :155  Lowres::init                 frameInitLowres + 4x extendPicBorder skipped (the GPU writes the planes)
  picyuv.cpp
    :51   PicYuv::create / destroy     planes in pinned memory (uploads are asynchronous DMA)
  weightPrediction.cpp
    :168  weightCost                   + x265glue_wp_cost first (one launch: weight applied on the fly + 8x8 SATDs + intra limit)
    :312,339,351  weightAnalyse        mcLuma / mcChroma skipped; + x265glue_wp_prepare before the sweep of a plane (motion
                                       compensation on the device), x265glue_wp_done before mcbuf is freed
"""
import os
import re
import sys


class Patch:
    def __init__(self, path):
        self.path = path
        self.lines = open(path).read().split("\n")
        self.inserts = []       # (index to insert BEFORE, text)
        self.replaces = {}      # index -> new text

    def find(self, pattern, start=0, nth=1):
        rx = re.compile(pattern)
        seen = 0
        for i in range(start, len(self.lines)):
            if rx.search(self.lines[i]):
                seen += 1
                if seen == nth:
                    return i
        raise SystemExit("make_gpu_sources: anchor not found in %s: %s" % (self.path, pattern))

    def before(self, idx, text):
        self.inserts.append((idx, text))

    def after(self, idx, text):
        self.inserts.append((idx + 1, text))

    def write(self, out):
        lines = list(self.lines)
        for idx, text in self.replaces.items():
            lines[idx] = text
        for idx, text in sorted(self.inserts, key=lambda t: -t[0]):
            lines.insert(idx, text)
        open(out, "w").write("\n".join(lines))


def slicetype(src, out):
    p = Patch(src)
    p.after(p.find(r'^#include "ratecontrol\.h"'), '#include "x265_glue.h"')

    # one context per encoder
    i = p.find(r'^bool Lookahead::create\(\)')
    p.after(p.find(r'm_scratch = X265_MALLOC\(int, m_tld\[0\]\.widthInCU\);', i), '    x265glue_open(this);')
    i = p.find(r'^void Lookahead::destroy\(\)')
    p.before(p.find(r'X265_FREE\(m_scratch\);', i), '    x265glue_close(this);')

    # pre-lookahead: the list in one call
    i = p.find(r'^void PreLookaheadGroup::processTasks\(int workerThreadID\)')
    j = p.find(r'^\s*m_lock\.acquire\(\);', i)
    p.after(j, '\n'.join([
        '    {',
        '        /* the first thread to arrive hands the whole list to the GPU (INTEGRATION.md 2); bonded peers find it taken */',
        '        int first = m_jobAcquired, count = m_jobTotal - m_jobAcquired;',
        '        m_jobAcquired = m_jobTotal;',
        '        m_lock.release();',
        '        if (count > 0)',
        '            x265glue_pre_list(&m_lookahead, m_preframes + first, count);',
        '        (void)tld;',
        '        return;',
        '    }']))

    # the lowres planes copied back for weightPrediction.cpp have landed before the frames leave the lookahead
    i = p.find(r'^void Lookahead::slicetypeDecide\(\)')
    j = p.find(r'dequeue all frames from inputQueue that are about to be enqueued', i)
    k = j
    while not re.search(r'm_inputLock\.acquire\(\);', p.lines[k]):
        k -= 1
    p.before(k, '    x265glue_sync(this);')

    # cuTree: memsets, pointer swaps, propagate steps, finish
    i = p.find(r'^void Lookahead::cuTree\(Lowres \*\*frames, int numframes, bool bIntra\)')
    j = p.find(r'^void Lookahead::estimateCUPropagate\(', i)
    rx = re.compile(r'^(\s*)memset\((frames\[\w+\])->propagateCost, 0, m_cuCount \* sizeof\(uint16_t\)\);')
    rs = re.compile(r'^(\s*)std::swap\((frames\[\w+\])->propagateCost, (frames\[\w+\])->propagateCost\);')
    nz = ns = 0
    for k in range(i, j):
        m = rx.match(p.lines[k])
        if m:
            p.after(k, '%sx265glue_ct_zero(this, %s);' % (m.group(1), m.group(2)))
            nz += 1
        m = rs.match(p.lines[k])
        if m:
            p.before(k, '%sx265glue_ct_preswap(this, %s, %s);' % m.groups())
            p.after(k, '%sx265glue_ct_postswap(this, %s, %s);' % m.groups())
            ns += 1
    if nz < 4 or ns != 2:
        raise SystemExit("make_gpu_sources: cuTree anchors: %d memsets, %d swaps" % (nz, ns))
    k = p.find(r'^\s*for \(uint16_t blocky = 0; blocky < m_8x8Height; blocky\+\+\)', j)
    p.before(k, '    if (x265glue_propagate(this, frames, averageDuration, p0, p1, b, referenced)) { } else')
    i = p.find(r'^void Lookahead::cuTreeFinish\(')
    p.after(p.find(r'^\{', i), '    if (x265glue_ct_finish(this, frame, averageDuration, ref0Distance)) return;')
    p.before(p.find(r'^\}', i), '    x265glue_ct_finished(this, frame, averageDuration, ref0Distance);')

    # a batch of estimates in one call
    i = p.find(r'^void CostEstimateGroup::finishBatch\(\)')
    p.after(p.find(r'^\{', i), '\n'.join([
        '    if (x265glue_finish_batch(&m_lookahead, m_frames, &m_estimates[0].p0, m_jobTotal))',
        '    {',
        '        m_jobTotal = m_jobAcquired = 0;',
        '        return;',
        '    }']))

    # one estimate: through the look-ahead estimate cache; afterwards the reference's cached branch returns it
    i = p.find(r'CostEstimateGroup::estimateFrameCost\(LookaheadTLD& tld')
    p.after(p.find(r'^\s*int64_t\s+score = 0;', i), '    x265glue_ensure(&m_lookahead, m_frames, p0, p1, b);')
    p.write(out)


def lowres(src, out):
    p = Patch(src)
    i = p.find(r'^using namespace X265_NS;')
    p.after(i, '\n'.join(['#include "x265_glue.h"', '#define x265_malloc x265glue_malloc', '#define x265_free x265glue_free']))
    i = p.find(r'^bool Lowres::create\(')
    p.after(p.find(r'isLowres = true;', i), '    x265glue_arena_begin(x265glue_lowres_bytes(origPic->m_picWidth, origPic->m_picHeight, origPic->m_lumaMarginX, '
                                            'origPic->m_lumaMarginY, _bframes, (int)sizeof(pixel)));')
    p.before(p.find(r'^\s*return true;', i), '    x265glue_arena_end();')
    p.after(p.find(r'^fail:', i), '    x265glue_arena_end();')
    i = p.find(r'^void Lowres::init\(')
    p.before(p.find(r'downscale and generate 4 hpel planes for lookahead', i), '    if (!x265glue_active())\n    {')
    p.before(p.find(r'fpelPlane\[0\] = lowresPlane\[0\];', i), '    }')
    p.write(out)


def picyuv(src, out):
    p = Patch(src)
    i = p.find(r'^using namespace X265_NS;')
    p.after(i, '#include "x265_glue.h"')
    i = p.find(r'^bool PicYuv::create\(')
    p.before(i, '#define x265_malloc x265glue_malloc\n#define x265_free x265glue_free')
    p.before(p.find(r'the first picture allocated by the encoder will be asked to generate these', i), '#undef x265_malloc\n#undef x265_free')
    i = p.find(r'^void PicYuv::destroy\(\)')
    p.before(i, '#define x265_malloc x265glue_malloc\n#define x265_free x265glue_free')
    p.after(p.find(r'^\}', i), '#undef x265_malloc\n#undef x265_free')
    p.write(out)


def weightpred(src, out):
    """encoder/weightPrediction.cpp: weightAnalyse keeps its float guesses, its sweep and its decisions; the three pixel
    loops (mcLuma, mcChroma, weightCost) run on the planes the lookahead left on the device"""
    p = Patch(src)
    p.after(p.find(r'^using namespace X265_NS;'), '#include "x265_glue.h"')
    i = p.find(r'^uint32_t weightCost\(pixel \*\s+fenc,')
    p.after(p.find(r'^\{', i), '\n'.join([
        '    {',
        '        uint32_t gpuCost;',
        '        if (x265glue_wp_cost(w != NULL, w ? w->inputWeight : 0, w ? (int)w->log2WeightDenom : 0, w ? w->inputOffset : 0, &gpuCost))',
        '            return gpuCost;',
        '    }']))
    i = p.find(r'^void weightAnalyse\(Slice& slice, Frame& frame, x265_param& param\)')
    k = p.find(r'^\s*mcLuma\(mcbuf, refLowres, mvs\);', i)
    p.replaces[k] = p.lines[k].replace('mcLuma(', 'if (!x265glue_active()) mcLuma(')
    n = 0
    for k in range(i, len(p.lines)):
        if re.match(r'^\s*mcChroma\(mcbuf, fref, stride, mvs, cache, height, width\);', p.lines[k]):
            p.replaces[k] = p.lines[k].replace('mcChroma(', 'if (!x265glue_active()) mcChroma(')
            n += 1
        if re.match(r'^\s*X265_FREE\(mcbuf\);', p.lines[k]):
            p.before(k, re.match(r'^(\s*)', p.lines[k]).group(1) + 'x265glue_wp_done();')
    if n != 2:
        raise SystemExit("make_gpu_sources: mcChroma anchors: %d" % n)
    k = p.find(r'^\s*uint32_t origscore = weightCost\(orig, fref, weightTemp, stride, cache, width, height, NULL, !plane\);', i)
    p.before(k, '            x265glue_wp_prepare(&frame, refFrame, plane, mvs);')
    p.write(out)


def main():
    ref, outdir = sys.argv[1], sys.argv[2]
    slicetype(os.path.join(ref, "encoder/slicetype.cpp"), os.path.join(outdir, "slicetype_gpu.cpp"))
    lowres(os.path.join(ref, "common/lowres.cpp"), os.path.join(outdir, "lowres_gpu.cpp"))
    picyuv(os.path.join(ref, "common/picyuv.cpp"), os.path.join(outdir, "picyuv_gpu.cpp"))
    weightpred(os.path.join(ref, "encoder/weightPrediction.cpp"), os.path.join(outdir, "weightpred_gpu.cpp"))


if __name__ == "__main__":
    main()
