/* la_core.h -- warp-uniform control logic of the lookahead CU estimate, as a small state machine.
 *
 * Product code (part of libx265cu.so).  Plain scalar C++ usable from host and device: the CUDA
 * kernels (x265cu_estimate.cuh) run it redundantly in every lane of the warp that owns a CU (so
 * control flow stays warp-uniform) while the pixel work of each pass -- up to 8 candidate blocks
 * -- is spread over the 32 lanes (one 4x4 sub-block per lane, one candidate per quad).
 * tests/core_emul.cpp compiles the very same header on the CPU with scalar evaluators to prove
 * the state machine against the oracle before any GPU time is spent.
 *
 * What it restates (file:line in /root/reference/x265_1.9/source):
 *   MV candidates / MVP selection / skip shortcut   encoder/slicetype.cpp:2117-2159
 *   MotionEstimate::motionEstimate, lowres HEX path  encoder/motion.cpp:587-624,670-742,1081-1119
 *   list / bidir / intra decision and accumulation   encoder/slicetype.cpp:2161-2224
 * Every comparison is a strict '<' taken in the reference's candidate order ("first minimum wins").
 */
#ifndef X265CU_LA_CORE_H
#define X265CU_LA_CORE_H

#include <stdint.h>

#if defined(__CUDACC__)
#define LA_HD __host__ __device__ __forceinline__
#else
#define LA_HD static inline
#endif

#define LA_COST_MAX (1 << 28)
#define LA_LOWRES_COST_MASK ((1 << 14) - 1)
#define LA_LOWRES_COST_SHIFT 14
#define LA_MERANGE 16

/* one candidate block of a pass */
struct LaCand
{
    int valid;
    int qx, qy;     /* quarter-pel MV relative to the CU position */
    int satd;       /* 1: SATD, 0: SAD */
    int addMv;      /* add mvcost(qx,qy) to the measured distortion */
};

enum LaPhase { LA_PH_CAND = 0, LA_PH_START, LA_PH_HEX6, LA_PH_HEX3, LA_PH_SQ8, LA_PH_HPEL, LA_PH_QPEL, LA_PH_DONE };

struct LaSearch
{
    int phase;
    int minx, miny, maxx, maxy;   /* full-pel search bounds (mvmin/mvmax) */
    int numc;
    int c0, c1, c2, c3;           /* packed candidate MVs: (uint16)x | y << 16 */
    int bidir;
    int mvpx, mvpy;               /* unclipped qpel MVP (setMVP) */
    int skipCost;
    int pmx, pmy, bprecost;       /* clipped qpel MVP and its SAD */
    int bmx, bmy, bcost;          /* running best (full-pel until the square refine, then qpel) */
    int dir, iter;
    int outx, outy, outcost;
};

LA_HD int la_pack_mv(int x, int y) { return (int)(((uint32_t)x & 0xffffu) | ((uint32_t)y << 16)); }
LA_HD int la_mv_x(int p) { return (int)(int16_t)(p & 0xffff); }
LA_HD int la_mv_y(int p) { return (int)(int16_t)((uint32_t)p >> 16); }
LA_HD int la_clip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }

/* radius-2 hexagon with wrap-around copies, (x-1)%6 table and the unit square (motion.cpp:64-66),
 * packed as 4-bit fields (value + 2) so that lane-dependent lookups need no memory */
LA_HD int la_hex2x(int i) { return (int)((0x01343101u >> (4 * i)) & 15) - 2; }   /* -1,-2,-1,1,2,1,-1,-2 */
LA_HD int la_hex2y(int i) { return (int)((0x20024420u >> (4 * i)) & 15) - 2; }   /* -2,0,2,2,0,-2,-2,0   */
LA_HD int la_mod6m1(int i) { return (int)((0x05432105u >> (4 * i)) & 15); }       /* 5,0,1,2,3,4,5,0      */
/* square1[0..8] = (0,0),(0,-1),(0,1),(-1,0),(1,0),(-1,-1),(-1,1),(1,-1),(1,1), stored as value + 1 */
LA_HD int la_sq1x(int i) { return (int)((0x220020111ull >> (4 * i)) & 15) - 1; }
LA_HD int la_sq1y(int i) { return (int)((0x202011201ull >> (4 * i)) & 15) - 1; }

/* mvcost(mv) = LUT[mv.x - mvp.x] + LUT[mv.y - mvp.y], returned as uint16_t (bitcost.h:42-45) */
LA_HD int la_mvcost(const uint16_t* lut, const LaSearch& s, int qx, int qy)
{
    return (int)(uint16_t)(lut[qx - s.mvpx] + lut[qy - s.mvpy]);
}

LA_HD bool la_in_range(const LaSearch& s) { return s.bmx >= s.minx && s.bmx <= s.maxx && s.bmy >= s.miny && s.bmy <= s.maxy; }

/* clip the MVP to the qpel search bounds (motion.cpp:600-601) and enter the START pass */
LA_HD void la_enter_start(LaSearch& s)
{
    s.pmx = la_clip(s.minx * 4, s.maxx * 4, s.mvpx);
    s.pmy = la_clip(s.miny * 4, s.maxy * 4, s.mvpy);
    s.phase = LA_PH_START;
}

/* begin the search of one list for CU (cuX, cuY).  nb[] = packed MVs of the right, below,
 * below-left, below-right neighbours in that order, already filtered by availability
 * (slicetype.cpp:2117-2128). */
LA_HD void la_search_begin(LaSearch& s, int cuX, int cuY, int wCU, int hCU, int bidir, int numc, int n0, int n1, int n2, int n3)
{
    s.minx = -cuX * 8 - 8;
    s.miny = -cuY * 8 - 8;
    s.maxx = (wCU - cuX - 1) * 8 + 8;
    s.maxy = (hCU - cuY - 1) * 8 + 8;
    s.bidir = bidir;
    s.numc = numc;
    s.c0 = n0; s.c1 = n1; s.c2 = n2; s.c3 = n3;
    s.mvpx = s.mvpy = 0;
    s.skipCost = 0x7fffffff;
    s.bprecost = s.bcost = LA_COST_MAX;
    s.bmx = s.bmy = s.dir = s.iter = 0;
    s.outx = s.outy = 0;
    s.outcost = LA_COST_MAX;
    if (numc)
        s.phase = LA_PH_CAND;
    else
        la_enter_start(s);
}

/* q-th candidate (0..7) of the current pass */
LA_HD LaCand la_candidate(const LaSearch& s, int q)
{
    LaCand c;
    c.valid = 0; c.qx = c.qy = 0; c.satd = 0; c.addMv = 1;
    switch (s.phase)
    {
    case LA_PH_CAND:
    {
        int p = q == 0 ? s.c0 : (q == 1 ? s.c1 : (q == 2 ? s.c2 : s.c3));
        c.valid = q < s.numc;
        c.qx = la_mv_x(p); c.qy = la_mv_y(p);
        c.satd = 1; c.addMv = 0;
        break;
    }
    case LA_PH_START:
        if (q == 0) { c.valid = 1; c.qx = s.pmx; c.qy = s.pmy; c.addMv = 0; }
        else if (q == 1) { c.valid = ((s.pmx | s.pmy) & 3) != 0; c.qx = ((s.pmx + 2) >> 2) * 4; c.qy = ((s.pmy + 2) >> 2) * 4; }
        else if (q == 2) { c.valid = (s.pmx | s.pmy) != 0; }
        break;
    case LA_PH_HEX6:
        c.valid = q < 6;
        c.qx = (s.bmx + la_hex2x((q + 1) & 7)) * 4; c.qy = (s.bmy + la_hex2y((q + 1) & 7)) * 4;
        break;
    case LA_PH_HEX3:
        c.valid = q < 3;
        c.qx = (s.bmx + la_hex2x((s.dir + q) & 7)) * 4; c.qy = (s.bmy + la_hex2y((s.dir + q) & 7)) * 4;
        break;
    case LA_PH_SQ8:
        c.valid = 1;
        c.qx = (s.bmx + la_sq1x(q + 1)) * 4; c.qy = (s.bmy + la_sq1y(q + 1)) * 4;
        break;
    case LA_PH_HPEL:
        c.valid = q < 4;
        c.qx = s.bmx + la_sq1x((q + 1) & 7) * 2; c.qy = s.bmy + la_sq1y((q + 1) & 7) * 2;
        break;
    case LA_PH_QPEL:
        c.valid = q < 5;
        c.qx = s.bmx + la_sq1x(q & 7); c.qy = s.bmy + la_sq1y(q & 7);
        c.satd = 1;
        break;
    default:
        break;
    }
    return c;
}

/* consume the costs of the current pass (cost[k] already includes mvcost where addMv was set;
 * entries of invalid candidates are ignored) and move to the next phase */
LA_HD void la_update(LaSearch& s, const int cost[8], const uint16_t* lut)
{
    switch (s.phase)
    {
    case LA_PH_CAND:
    {
        int mvpcost = LA_COST_MAX;
        for (int k = 0; k < 4; k++)
        {
            if (k >= s.numc) break;
            int p = k == 0 ? s.c0 : (k == 1 ? s.c1 : (k == 2 ? s.c2 : s.c3));
            if (cost[k] < mvpcost) { mvpcost = cost[k]; s.mvpx = la_mv_x(p); s.mvpy = la_mv_y(p); }
            /* holds the cost of the last candidate measured while the best MVP is still zero */
            if (!(s.mvpx | s.mvpy) && s.bidir)
                s.skipCost = cost[k];
        }
        la_enter_start(s);
        break;
    }
    case LA_PH_START:
        s.bprecost = cost[0];
        s.bmx = (s.pmx + 2) >> 2; s.bmy = (s.pmy + 2) >> 2;
        s.bcost = ((s.pmx | s.pmy) & 3) ? cost[1] : cost[0];
        if ((s.pmx | s.pmy) && cost[2] < s.bcost) { s.bcost = cost[2]; s.bmx = s.bmy = 0; }
        s.phase = LA_PH_HEX6;
        break;
    case LA_PH_HEX6:
    {
        int best = -1;
        for (int k = 0; k < 6; k++)
            if (cost[k] < s.bcost) { s.bcost = cost[k]; best = k; }
        s.phase = LA_PH_SQ8;
        if (best >= 0)
        {
            s.dir = best;
            s.bmx += la_hex2x(best + 1); s.bmy += la_hex2y(best + 1);
            s.iter = (LA_MERANGE >> 1) - 1;
            if (la_in_range(s)) s.phase = LA_PH_HEX3;
        }
        break;
    }
    case LA_PH_HEX3:
    {
        int step = -1;
        for (int k = 0; k < 3; k++)
            if (cost[k] < s.bcost) { s.bcost = cost[k]; step = k; }
        s.phase = LA_PH_SQ8;
        if (step >= 0)
        {
            s.dir = la_mod6m1(s.dir + step);      /* mod6m1[dir + (step + 1) - 2 + 1] */
            s.bmx += la_hex2x(s.dir + 1); s.bmy += la_hex2y(s.dir + 1);
            s.iter--;
            if (s.iter > 0 && la_in_range(s)) s.phase = LA_PH_HEX3;
        }
        break;
    }
    case LA_PH_SQ8:
    {
        int best = 0;
        for (int k = 0; k < 8; k++)
            if (cost[k] < s.bcost) { s.bcost = cost[k]; best = k + 1; }
        s.bmx += la_sq1x(best); s.bmy += la_sq1y(best);
        if (s.bprecost < s.bcost) { s.bmx = s.pmx; s.bmy = s.pmy; s.bcost = s.bprecost; }
        else { s.bmx *= 4; s.bmy *= 4; }
        if (!s.bcost)
        {
            s.outcost = la_mvcost(lut, s, s.bmx, s.bmy);
            s.outx = s.bmx; s.outy = s.bmy;
            s.phase = LA_PH_DONE;
        }
        else
            s.phase = LA_PH_HPEL;
        break;
    }
    case LA_PH_HPEL:
    {
        int bdir = 0;
        for (int k = 0; k < 4; k++)
            if (cost[k] < s.bcost) { s.bcost = cost[k]; bdir = k + 1; }
        s.bmx += la_sq1x(bdir) * 2; s.bmy += la_sq1y(bdir) * 2;
        s.phase = LA_PH_QPEL;
        break;
    }
    case LA_PH_QPEL:
    {
        int bdir = 0;
        s.bcost = cost[0];       /* SATD re-measure at the half-pel winner (may go up) */
        for (int k = 1; k < 5; k++)
            if (cost[k] < s.bcost) { s.bcost = cost[k]; bdir = k; }
        s.bmx += la_sq1x(bdir); s.bmy += la_sq1y(bdir);
        s.outcost = s.bcost; s.outx = s.bmx; s.outy = s.bmy;
        s.phase = LA_PH_DONE;
        break;
    }
    default:
        break;
    }
    if (s.phase == LA_PH_DONE && s.skipCost < 64 && s.skipCost < s.outcost && s.bidir)
    {
        /* bidir-only zero-MV skip shortcut (slicetype.cpp:2155-2159) */
        s.outcost = s.skipCost;
        s.outx = s.outy = 0;
    }
}

/* ---- per-CU decision after the list costs are known (slicetype.cpp:2161-2224) ---- */
struct LaCuResult
{
    int bcost, bcostAq, listused, scored, intraMb;
    uint16_t lowresCost;
};

/* listCost[i] < 0 means list i is not part of this estimate.  bicost0/bicost1: SATD of the
 * avg(L0-MC, L1-MC) and co-located average candidates (B only). */
LA_HD LaCuResult la_cu_finish(int cuX, int cuY, int wCU, int hCU, int bidir, int cost0, int cost1,
                              int bicost0, int bicost1, int intraCost, int hasInvQ, int invQ)
{
    LaCuResult r;
    int bcost = LA_COST_MAX, listused = 0;
    if (cost0 < bcost) { bcost = cost0; listused = 1; }
    if (bidir)
    {
        if (cost1 < bcost) { bcost = cost1; listused = 2; }
        if (bicost0 < bcost) { bcost = bicost0; listused = 3; }
        if (bicost1 < bcost) { bcost = bicost1; listused = 3; }
        bcost += 4;
    }
    else
    {
        bcost += 4;
        if (intraCost < bcost) { bcost = intraCost; listused = 0; }
    }
    r.scored = (cuX > 0 && cuX < wCU - 1 && cuY > 0 && cuY < hCU - 1) || wCU <= 2 || hCU <= 2;
    r.bcost = bcost;
    r.bcostAq = (r.scored && hasInvQ) ? ((bcost * invQ + 128) >> 8) : bcost;
    r.listused = listused;
    r.intraMb = (!listused && !bidir) ? 1 : 0;
    int capped = bcost < LA_LOWRES_COST_MASK ? bcost : LA_LOWRES_COST_MASK;
    r.lowresCost = (uint16_t)(capped | (listused << LA_LOWRES_COST_SHIFT));
    return r;
}

/* the two sources of a quarter-pel reference block (lowres.h:62-103): hpel plane index and the
 * full-pel offset of each; second source only when the MV has an odd component */
struct LaMcSrc { int planeA, ax, ay, avg, planeB, bx, by; };

LA_HD LaMcSrc la_mc_src(int qx, int qy)
{
    LaMcSrc m;
    m.planeA = (qy & 2) | ((qx & 2) >> 1);
    m.ax = qx >> 2; m.ay = qy >> 2;
    m.avg = (qx | qy) & 1;
    int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
    m.planeB = (qy2 & 2) | ((qx2 & 2) >> 1);
    m.bx = qx2 >> 2; m.by = qy2 >> 2;
    return m;
}

#endif /* X265CU_LA_CORE_H */
