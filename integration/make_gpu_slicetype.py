#!/usr/bin/env python3
"""Make a temporary GPU-hooked copy of the reference's encoder/slicetype.cpp (INTEGRATION PROOF).

usage: make_gpu_slicetype.py <reference slicetype.cpp> <out.cpp>

Inserts the call-outs of integration/x265_glue.h at the sites INTEGRATION.md names; holds no
reference source (one-line anchors only).  The copy lives under oracle/_ref/ (git-ignored).
  :851   PreLookaheadGroup::processTasks  -> x265glue_pre        (GPU lowres planes + intra estimate)
  :486   weightsAnalyse accepted a weight  -> x265glue_weight
  :2007  estimateFrameCost, before the CPU loops -> `if (x265glue_estimate(...)) {} else <CPU loops>`
  :1760  estimateCUPropagate, before its CU loops -> `if (x265glue_propagate(...)) {} else <CPU loops>` (cuTree)
"""
import re
import sys


def main():
    src_path, out_path = sys.argv[1], sys.argv[2]
    lines = open(src_path).read().split("\n")

    def find(pattern, start=0):
        rx = re.compile(pattern)
        for i in range(start, len(lines)):
            if rx.search(lines[i]):
                return i
        raise SystemExit("make_gpu_slicetype: anchor not found: %s" % pattern)

    inserts = []
    i = find(r'^#include "ratecontrol\.h"')
    inserts.append((i + 1, '#include "x265_glue.h"'))
    i = find(r'preFrame->m_lowresInit = true;')
    inserts.append((i, '        x265glue_pre(&m_lookahead, preFrame);'))
    i = find(r'^\s*weightedRef\.isWeighted = true;')
    inserts.append((i + 1, '        x265glue_weight(wp.inputWeight, (int)wp.log2WeightDenom, wp.inputOffset);'))
    i = find(r'CostEstimateGroup::estimateFrameCost\(LookaheadTLD& tld')
    j = find(r'^\s*if \(!m_batchMode && m_lookahead\.m_numCoopSlices > 1', i)
    inserts.append((j, '        if (x265glue_estimate(&m_lookahead, m_frames, p0, p1, b, bDoSearch, m_batchMode)) { } else'))
    i = find(r'^void Lookahead::estimateCUPropagate\(')
    j = find(r'^\s*for \(uint16_t blocky = 0; blocky < m_8x8Height; blocky\+\+\)', i)
    inserts.append((j, '    if (x265glue_propagate(this, frames, fpsFactor, bipredWeight, p0, p1, b, referenced)) { } else'))
    for idx, text in sorted(inserts, key=lambda t: -t[0]):
        lines.insert(idx, text)
    open(out_path, "w").write("\n".join(lines))


if __name__ == "__main__":
    main()
