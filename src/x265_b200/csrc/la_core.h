/* la_core.h -- warp-uniform control logic of the lookahead CU estimate, phase by phase.
 *
 * Product code (part of libx265cu.so).  Plain scalar C++ usable from host and device: the CUDA
 * search kernel (x265cu_kernels.cuh) runs it redundantly in every lane of the warp that owns a CU
 * row (so control flow stays warp-uniform) while the pixel work of each pass -- up to 8 candidate
 * blocks -- is spread over the 32 lanes (one 4x4 sub-block per lane, one candidate per quad).
 * tests/core_emul.cpp compiles the very same header on the CPU with scalar evaluators to prove
 * the logic against the reference's golden traces before any GPU time is spent.
 *
 * What it restates (file:line in /root/reference/x265_1.9/source):
 *   MV candidates / MVP selection / skip shortcut   encoder/slicetype.cpp:2117-2159
 *   MotionEstimate::motionEstimate, lowres HEX path  encoder/motion.cpp:587-624,670-742,1081-1119
 *   list / bidir / intra decision and accumulation   encoder/slicetype.cpp:2161-2224
 *
 * Passes of one list search (each pass = candidates measured in parallel, then one update):
 *   CAND  <=4 neighbour MVs, SATD, no mvcost          -> MVP (+ skipCost)
 *   START qpel MVP (SAD, no mvcost), rounded MVP, zero  -> search start
 *   HEX6  6-point hexagon radius 2                      -> move or stay
 *   HEX3  half hexagon, up to merange/2-1 = 7 times     -> move or stop
 *   SQ8   8-point unit square                           -> full-pel winner (or the qpel MVP)
 *   HPEL  4 half-pel SADs                               -> half-pel winner
 *   QPEL  SATD re-measure + 4 quarter-pel SATDs         -> result
 * A pass whose winner is decided by "first strict minimum in candidate order" is fed with the
 * minimum of the packed keys (cost << 3 | k) over its valid candidates, which is exactly how the
 * reference packs direction bits into the cost (motion.cpp:693-725): one warp min-reduction
 * replaces the sequential COPYn_IF_LT chain bit-exactly.
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
#define LA_KEY_NONE 0xFFFFFFFFu
#define LA_LOWRES_COST_MASK ((1 << 14) - 1)
#define LA_LOWRES_COST_SHIFT 14
#define LA_MERANGE 16

struct LaSearch
{
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
LA_HD int la_mv_y(int p) { return p >> 16; }
LA_HD int la_clip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
LA_HD uint32_t la_key(int cost, int k) { return ((uint32_t)cost << 3) | (uint32_t)k; }

/* radius-2 hexagon with wrap-around copies, (x-1)%6 table and the unit square (motion.cpp:64-66),
 * packed as 4-bit fields so that lane-dependent lookups need no memory */
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

/* begin the search of one list for CU (cuX, cuY).  n0..n3 = packed MVs of the right, below,
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
    s.pmx = s.pmy = 0;
    s.bprecost = s.bcost = LA_COST_MAX;
    s.bmx = s.bmy = s.dir = s.iter = 0;
    s.outx = s.outy = 0;
    s.outcost = LA_COST_MAX;
}

/* ---- CAND: the q-th neighbour MV (q < numc), measured with SATD, no mvcost ---- */
LA_HD int la_cand_mv(const LaSearch& s, int q) { return q == 0 ? s.c0 : (q == 1 ? s.c1 : (q == 2 ? s.c2 : s.c3)); }

LA_HD void la_upd_cand(LaSearch& s, int c0, int c1, int c2, int c3)
{
    int mvpcost = LA_COST_MAX;
    for (int k = 0; k < 4; k++)
    {
        if (k >= s.numc) break;
        int cost = k == 0 ? c0 : (k == 1 ? c1 : (k == 2 ? c2 : c3));
        int p = la_cand_mv(s, k);
        if (cost < mvpcost) { mvpcost = cost; s.mvpx = la_mv_x(p); s.mvpy = la_mv_y(p); }
        /* holds the cost of the last candidate measured while the best MVP is still zero */
        if (!(s.mvpx | s.mvpy) && s.bidir)
            s.skipCost = cost;
    }
}

/* ---- START (motion.cpp:600-624): q0 = clipped qpel MVP (no mvcost), q1 = rounded MVP (only when
 * the MVP is sub-pel), q2 = zero MV (only when the MVP is non-zero) ---- */
LA_HD void la_enter_start(LaSearch& s)
{
    s.pmx = la_clip(s.minx * 4, s.maxx * 4, s.mvpx);
    s.pmy = la_clip(s.miny * 4, s.maxy * 4, s.mvpy);
}
LA_HD bool la_start_subpel(const LaSearch& s) { return ((s.pmx | s.pmy) & 3) != 0; }
LA_HD bool la_start_nonzero(const LaSearch& s) { return (s.pmx | s.pmy) != 0; }

LA_HD void la_upd_start(LaSearch& s, int c0, int c1, int c2)
{
    s.bprecost = c0;
    s.bmx = (s.pmx + 2) >> 2; s.bmy = (s.pmy + 2) >> 2;
    s.bcost = la_start_subpel(s) ? c1 : c0;
    if (la_start_nonzero(s) && c2 < s.bcost) { s.bcost = c2; s.bmx = s.bmy = 0; }
}

/* ---- HEX6: candidate q < 6 at bm + hex2[q + 1]; returns true when the half-hexagon loop runs ---- */
LA_HD bool la_upd_hex6(LaSearch& s, uint32_t key)
{
    if (key != LA_KEY_NONE && (int)(key >> 3) < s.bcost)
    {
        int best = (int)(key & 7);
        s.bcost = (int)(key >> 3);
        s.dir = best;
        s.bmx += la_hex2x(best + 1); s.bmy += la_hex2y(best + 1);
        s.iter = (LA_MERANGE >> 1) - 1;
        return la_in_range(s);
    }
    return false;
}

/* ---- HEX3: candidate q < 3 at bm + hex2[dir + q]; returns true to run another round ---- */
LA_HD bool la_upd_hex3(LaSearch& s, uint32_t key)
{
    if (key != LA_KEY_NONE && (int)(key >> 3) < s.bcost)
    {
        int step = (int)(key & 7);
        s.bcost = (int)(key >> 3);
        s.dir = la_mod6m1(s.dir + step);      /* mod6m1[dir + (step + 1) - 2 + 1] */
        s.bmx += la_hex2x(s.dir + 1); s.bmy += la_hex2y(s.dir + 1);
        s.iter--;
        return s.iter > 0 && la_in_range(s);
    }
    return false;
}

/* ---- SQ8: candidate q < 8 at bm + square1[q + 1]; afterwards the winner becomes quarter-pel.
 * Returns true when the sub-pel refine runs, false when the search ends here (zero residual) ---- */
LA_HD bool la_upd_sq8(LaSearch& s, uint32_t key, const uint16_t* lut)
{
    if (key != LA_KEY_NONE && (int)(key >> 3) < s.bcost)
    {
        int best = (int)(key & 7) + 1;
        s.bcost = (int)(key >> 3);
        s.bmx += la_sq1x(best); s.bmy += la_sq1y(best);
    }
    if (s.bprecost < s.bcost) { s.bmx = s.pmx; s.bmy = s.pmy; s.bcost = s.bprecost; }
    else { s.bmx *= 4; s.bmy *= 4; }
    if (!s.bcost)
    {
        s.outcost = la_mvcost(lut, s, s.bmx, s.bmy);
        s.outx = s.bmx; s.outy = s.bmy;
        return false;
    }
    return true;
}

/* ---- HPEL: candidate q < 4 at bm + square1[q + 1] * 2 (quarter-pel units), SAD + mvcost ---- */
LA_HD void la_upd_hpel(LaSearch& s, uint32_t key)
{
    if (key != LA_KEY_NONE && (int)(key >> 3) < s.bcost)
    {
        int bdir = (int)(key & 7) + 1;
        s.bcost = (int)(key >> 3);
        s.bmx += la_sq1x(bdir) * 2; s.bmy += la_sq1y(bdir) * 2;
    }
}

/* ---- QPEL: q0 = SATD re-measure at bm (may go up), q1..q4 at bm + square1[q]; key covers q >= 1 ---- */
LA_HD void la_upd_qpel(LaSearch& s, int c0, uint32_t key)
{
    s.bcost = c0;
    if (key != LA_KEY_NONE && (int)(key >> 3) < s.bcost)
    {
        int bdir = (int)(key & 7);
        s.bcost = (int)(key >> 3);
        s.bmx += la_sq1x(bdir); s.bmy += la_sq1y(bdir);
    }
    s.outcost = s.bcost; s.outx = s.bmx; s.outy = s.bmy;
}

/* ---- speculative fast path ------------------------------------------------------------------
 * When nothing moves, the search visits positions that are all known once the MVP is known:
 *   START   qpel MVP pm, rounded MVP bm0 = (pm + 2) >> 2, zero
 *   HEX6    bm0 + hex2[1..6]          SQ8   bm0 + square1[1..8]
 *   HPEL    pm + 2 * square1[1..4]    QPEL  pm, pm + square1[1..4]
 * (the sub-pel refine is centred on pm both when the qpel MVP beats the full-pel winner and when pm
 * is full-pel and nothing moved).  The kernel measures all of them at once and feeds the costs to
 * la_fast_path(), which replays the reference's decisions in order and stops at the first step
 * whose speculative inputs do not apply (the search moved); it returns the stage at which the
 * one-pass-at-a-time evaluation has to resume, with `s` exactly as that stage expects it.  A cost
 * is a function of the position alone, so consuming it early is bit-exact. */
enum { LA_RESUME_DONE = 0, LA_RESUME_HEX6, LA_RESUME_HEX3, LA_RESUME_SQ8, LA_RESUME_HPEL, LA_RESUME_QPEL };

LA_HD int la_fast_path(LaSearch& s, int c0, int c1, int c2, uint32_t hexKey, uint32_t sqKey, uint32_t hpelKey,
                       int qc0, uint32_t qpelKey, const uint16_t* lut)
{
    la_upd_start(s, c0, c1, c2);
    if (s.bmx != ((s.pmx + 2) >> 2) || s.bmy != ((s.pmy + 2) >> 2))
        return LA_RESUME_HEX6;                          /* the zero MV won the start: hexagon is elsewhere */
    if (hexKey != LA_KEY_NONE && (int)(hexKey >> 3) < s.bcost)
        return la_upd_hex6(s, hexKey) ? LA_RESUME_HEX3 : LA_RESUME_SQ8;
    if (!la_upd_sq8(s, sqKey, lut))
        return LA_RESUME_DONE;                          /* zero residual: result is final */
    if (s.bmx != s.pmx || s.bmy != s.pmy)
        return LA_RESUME_HPEL;                          /* the square moved and the qpel MVP lost */
    if (hpelKey != LA_KEY_NONE && (int)(hpelKey >> 3) < s.bcost)
    {
        la_upd_hpel(s, hpelKey);
        return LA_RESUME_QPEL;
    }
    la_upd_qpel(s, qc0, qpelKey);
    return LA_RESUME_DONE;
}

/* bidir-only zero-MV skip shortcut (slicetype.cpp:2155-2159), applied once the search is done */
LA_HD void la_finish_skip(LaSearch& s)
{
    if (s.skipCost < 64 && s.skipCost < s.outcost && s.bidir)
    {
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

/* bicost0/bicost1: SATD of the avg(L0-MC, L1-MC) and co-located average candidates (B only). */
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
