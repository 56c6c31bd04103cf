#!/usr/bin/env python3
"""Make a temporary *hooked* copy of the reference's encoder/slicetype.cpp (TEST INFRASTRUCTURE).

usage: make_hooked_slicetype.py <reference slicetype.cpp> <out.cpp>

The copy lives only under oracle/_ref/ (git-ignored).  This script holds no reference source: it
finds a handful of one-line anchors in the file it is given and inserts call-outs to the hooks
declared in oracle/ref_hooks.h.  The call-outs only OBSERVE (trace which estimates the lookahead
ran, checksum their outputs); no arithmetic or control flow of the reference is changed: the
hooks are plain function calls that only read the reference's state (oracle/ref_shim.cpp).  The
stock CLI (oracle/_ref/x265_ref<d>) is linked with the UN-hooked object.

Hook sites (reference file:line in x265_1.9/source/encoder/slicetype.cpp):
  :851  after PreLookaheadGroup::processTasks finished a frame (m_lowresInit = true)
  :1919 CostEstimateGroup::finishBatch() entry, :1925 before its job counters are reset
  :2053 in estimateFrameCost(), after a non-cached estimate has been computed, before the
        B-frame scaling of the score
  :1668-1701 after each memset of a frame's propagateCost in Lookahead::cuTree
  :1839 end of Lookahead::estimateCUPropagate (before its optional cuTreeFinish), :1862 end of cuTreeFinish
"""
import re
import sys


def main():
    src_path, out_path = sys.argv[1], sys.argv[2]
    lines = open(src_path).read().split("\n")

    def find(pattern, start=0, nth=1):
        rx = re.compile(pattern)
        seen = 0
        for i in range(start, len(lines)):
            if rx.search(lines[i]):
                seen += 1
                if seen == nth:
                    return i
        raise SystemExit("make_hooked_slicetype: anchor not found: %s" % pattern)

    inserts = []  # (line index to insert BEFORE, text)

    # include the hook declarations after the last project include
    i = find(r'^#include "ratecontrol\.h"')
    inserts.append((i + 1, '#include "ref_hooks.h"'))

    # pre-lookahead done for one frame
    i = find(r'preFrame->m_lowresInit = true;')
    inserts.append((i + 1, '        x265ref_hook_pre(preFrame);'))

    # weightsAnalyse accepted a weight (:486); wp holds the final (scale, denom, offset)
    i = find(r'^\s*weightedRef\.isWeighted = true;')
    inserts.append((i + 1, '        x265ref_hook_weight(fenc.frameNum, ref.frameNum, wp.inputWeight, (int)wp.log2WeightDenom, wp.inputOffset);'))

    # batch begin / end
    i = find(r'^void CostEstimateGroup::finishBatch\(\)')
    j = find(r'^\{', i)
    inserts.append((j + 1, '    x265ref_hook_batch(1, m_jobTotal);'))
    k = find(r'm_jobTotal = m_jobAcquired = 0;', j)
    inserts.append((k, '    x265ref_hook_batch(0, m_jobTotal);'))

    # one estimate computed (non-cached branch)
    i = find(r'CostEstimateGroup::estimateFrameCost\(LookaheadTLD& tld')
    j = find(r'^\s*if \(b != p1\)\s*$', i)
    inserts.append((j, '        x265ref_hook_job(m_frames, p0, p1, b, bDoSearch[0], bDoSearch[1], m_batchMode, '
                       '(!m_batchMode && m_lookahead.m_numCoopSlices > 1 && ((p1 > b) || bDoSearch[0] || bDoSearch[1])));'))

    # cuTree: every memset of a propagateCost array, the end of each propagate step, the end of cuTreeFinish
    i = find(r'^void Lookahead::cuTree\(Lowres \*\*frames, int numframes, bool bIntra\)')
    j = find(r'^void Lookahead::estimateCUPropagate\(', i)
    rx = re.compile(r'^(\s*)memset\((frames\[\w+\])->propagateCost, 0, m_cuCount \* sizeof\(uint16_t\)\);')
    for k in range(i, j):
        m = rx.match(lines[k])
        if m:
            inserts.append((k + 1, '%sx265ref_hook_ctzero(%s);' % (m.group(1), m.group(2))))
    k = find(r'^\s*if \(m_param->rc\.vbvBufferSize && m_param->lookaheadDepth && referenced\)', j)
    inserts.append((k, '    x265ref_hook_propagate(frames, averageDuration, p0, p1, b, referenced);'))
    i = find(r'^void Lookahead::cuTreeFinish\(')
    k = find(r'^\}', i)
    inserts.append((k, '    x265ref_hook_ctfinish(frame, averageDuration, ref0Distance);'))

    for idx, text in sorted(inserts, key=lambda t: -t[0]):
        lines.insert(idx, text)
    open(out_path, "w").write("\n".join(lines))


if __name__ == "__main__":
    main()
