/* core_emul.cpp -- CPU emulation of the GPU search/cost kernels' control flow (TEST CODE).
 *
 * Compiles the product's warp-uniform state machine (src/x265_b200/csrc/la_core.h) on the host and
 * drives it with scalar evaluators taken from the oracle, one "pass" of up to 8 candidates at a
 * time exactly as the CUDA kernel does.  emul_estimate() has the signature of the oracle's
 * ola_estimate() so tests/test_core_emul.py can replay the reference traces through it: if the
 * state machine disagreed with the reference anywhere (tie-breaking, hexagon walk, skip shortcut,
 * bidir/intra decision...) the CRCs in the golden traces would catch it before GPU time is spent.
 */
#include "../oracle/x265la_oracle.h"
#include "../src/x265_b200/csrc/la_core.h"

#include <string.h>

namespace {

struct Refs
{
    pixel* plane[4];
    intptr_t stride;
};

int evalQ(const pixel* fenc, const Refs& r, intptr_t off, int qx, int qy, int satd)
{
    pixel blk[64];
    /* exercise la_mc_src (the kernel's address computation) instead of the oracle's own MC */
    LaMcSrc m = la_mc_src(qx, qy);
    const pixel* a = r.plane[m.planeA] + off + m.ax + (intptr_t)m.ay * r.stride;
    const pixel* b = r.plane[m.planeB] + off + m.bx + (intptr_t)m.by * r.stride;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            blk[8 * y + x] = m.avg ? (pixel)((a[y * r.stride + x] + b[y * r.stride + x] + 1) >> 1) : a[y * r.stride + x];
    return satd ? ola_satd8x8(fenc, 8, blk, 8) : ola_sad8x8(fenc, 8, blk, 8);
}

struct Acc { int64_t costEst, costEstAq; int intraMbs; };

/* search-shape statistics (design data for the speculative kernel; read with emul_stats()) */
enum { ST_TOTAL, ST_LASTROW, ST_MVP_IN_BELOW, ST_MVP_RIGHT_ONLY, ST_MVP_ZERO_NOCAND, ST_DISTINCT1, ST_DISTINCT2, ST_DISTINCT3,
       ST_HEX_MOVED, ST_HEX3_ROUNDS, ST_SQ_MOVED, ST_PRE_WON, ST_SUBPEL, ST_HPEL_MOVED, ST_QPEL_MOVED, ST_OUT_EQ_MVP,
       ST_LAST_EQ2, ST_LAST_TOTAL, ST_RIGHT_EQ_BELOWSET, ST_START_MOVED, ST_SKIP, ST_RIGHT_EQ2, ST_MVP_IN_BELOW_OR_R2,
       ST_RESUME0, ST_RESUME1, ST_RESUME2, ST_RESUME3, ST_RESUME4, ST_RESUME5, ST_N };
long long g_stats[ST_N];

void emulCU(ola_ctx* c, ola_frame* fenc, ola_frame* fref0, ola_frame* fref1, const Refs& wref0,
            int cuX, int cuY, int d0, int d1, const int doSearch[2], int lastRow, Acc& acc)
{
    const ola_geom& g = fenc->g;
    const int W = g.wCU, H = g.hCU;
    const int bidir = d1 > 0;
    const int cuXY = cuX + cuY * W;
    const intptr_t off = 8 * cuX + (intptr_t)8 * cuY * g.stride;
    pixel fb[64];
    for (int y = 0; y < 8; y++)
        memcpy(fb + 8 * y, fenc->plane[0] + off + (intptr_t)y * g.stride, 8 * sizeof(pixel));
    Refs r0 = { { fref0->plane[0], fref0->plane[1], fref0->plane[2], fref0->plane[3] }, g.stride };
    Refs r1 = { { fref1->plane[0], fref1->plane[1], fref1->plane[2], fref1->plane[3] }, g.stride };

    int listCost[2] = { LA_COST_MAX, LA_COST_MAX };
    for (int i = 0; i < 1 + bidir; i++)
    {
        int dist = (i ? d1 : d0) - 1;
        if (!doSearch[i])
        {
            listCost[i] = fenc->mvCosts[i][dist][cuXY];
            continue;
        }
        ola_mv* mv = fenc->mvs[i][dist];
        int nb[4] = { 0, 0, 0, 0 }, numc = 0;
        if (cuX < W - 1) nb[numc++] = la_pack_mv(mv[cuXY + 1].x, mv[cuXY + 1].y);
        if (!lastRow)
        {
            nb[numc++] = la_pack_mv(mv[cuXY + W].x, mv[cuXY + W].y);
            if (cuX > 0) nb[numc++] = la_pack_mv(mv[cuXY + W - 1].x, mv[cuXY + W - 1].y);
            if (cuX < W - 1) nb[numc++] = la_pack_mv(mv[cuXY + W + 1].x, mv[cuXY + W + 1].y);
        }
        const Refs& ref = i ? r1 : wref0;
        LaSearch s;
        la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb[0], nb[1], nb[2], nb[3]);
        g_stats[ST_TOTAL]++;
        if (lastRow) g_stats[ST_LASTROW]++;
        /* one pass = up to 8 candidates measured "in parallel", then one uniform update,
         * exactly the sequence of the CUDA search kernel */
        int cost[8];
        struct Pass
        {
            static uint32_t minKey(const int* c, const bool* valid, int from, int n)
            {
                uint32_t key = LA_KEY_NONE;
                for (int k = from; k < n; k++)
                    if (valid[k]) { uint32_t kk = la_key(c[k], k - (from ? 0 : 0)); if (kk < key) key = kk; }
                return key;
            }
        };
        bool valid[8];
        if (numc)
        {
            for (int q = 0; q < 4; q++)
            {
                cost[q] = LA_COST_MAX;
                if (q < numc)
                {
                    int p = la_cand_mv(s, q);
                    cost[q] = evalQ(fb, ref, off, la_mv_x(p), la_mv_y(p), 1);
                }
            }
            la_upd_cand(s, cost[0], cost[1], cost[2], cost[3]);
        }
        {
            int mvp = la_pack_mv(s.mvpx, s.mvpy);
            int r2 = (cuX < W - 2) ? la_pack_mv(mv[cuXY + 2].x, mv[cuXY + 2].y) : 0;
            if (!numc) g_stats[ST_MVP_ZERO_NOCAND]++;
            if (lastRow && numc) { g_stats[ST_LAST_TOTAL]++; if (mvp == r2) g_stats[ST_LAST_EQ2]++; }
            if (!lastRow)
            {
                int first = (cuX < W - 1) ? 1 : 0, inb = 0, dist = 0;
                for (int k = first; k < numc; k++)
                {
                    if (nb[k] == mvp) inb = 1;
                    int dup = 0;
                    for (int j = first; j < k; j++) dup |= nb[j] == nb[k];
                    dist += !dup;
                }
                if (inb) g_stats[ST_MVP_IN_BELOW]++; else g_stats[ST_MVP_RIGHT_ONLY]++;
                if (inb || mvp == r2) g_stats[ST_MVP_IN_BELOW_OR_R2]++;
                g_stats[ST_DISTINCT1 + dist - 1]++;
                if (first) { int e = 0; for (int k = 1; k < numc; k++) e |= nb[k] == nb[0]; if (e) g_stats[ST_RIGHT_EQ_BELOWSET]++; if (nb[0] == r2) g_stats[ST_RIGHT_EQ2]++; }
            }
        }
        la_enter_start(s);
        /* ---- speculative pass: every position of the no-move path measured up front, exactly the
         * candidate set of the CUDA kernel's search_mv() ---- */
        int resume;
        {
            const uint16_t* lut = c->mvcost;
            const int bm0x = (s.pmx + 2) >> 2, bm0y = (s.pmy + 2) >> 2;
            int c0 = evalQ(fb, ref, off, s.pmx, s.pmy, 0);
            int c1 = evalQ(fb, ref, off, bm0x * 4, bm0y * 4, 0) + la_mvcost(lut, s, bm0x * 4, bm0y * 4);
            int c2 = evalQ(fb, ref, off, 0, 0, 0) + la_mvcost(lut, s, 0, 0);
            uint32_t hexKey = LA_KEY_NONE, sqKey = LA_KEY_NONE, hpelKey = LA_KEY_NONE, qpelKey = LA_KEY_NONE;
            for (int q = 0; q < 6; q++)
            {
                int qx = (bm0x + la_hex2x(q + 1)) * 4, qy = (bm0y + la_hex2y(q + 1)) * 4;
                uint32_t k = la_key(evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(lut, s, qx, qy), q);
                if (k < hexKey) hexKey = k;
            }
            for (int q = 0; q < 8; q++)
            {
                int qx = (bm0x + la_sq1x(q + 1)) * 4, qy = (bm0y + la_sq1y(q + 1)) * 4;
                uint32_t k = la_key(evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(lut, s, qx, qy), q);
                if (k < sqKey) sqKey = k;
            }
            for (int q = 0; q < 4; q++)
            {
                int qx = s.pmx + la_sq1x(q + 1) * 2, qy = s.pmy + la_sq1y(q + 1) * 2;
                uint32_t k = la_key(evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(lut, s, qx, qy), q);
                if (k < hpelKey) hpelKey = k;
            }
            int qc0 = evalQ(fb, ref, off, s.pmx, s.pmy, 1) + la_mvcost(lut, s, s.pmx, s.pmy);
            for (int q = 1; q < 5; q++)
            {
                int qx = s.pmx + la_sq1x(q), qy = s.pmy + la_sq1y(q);
                uint32_t k = la_key(evalQ(fb, ref, off, qx, qy, 1) + la_mvcost(lut, s, qx, qy), q);
                if (k < qpelKey) qpelKey = k;
            }
            resume = la_fast_path(s, c0, c1, c2, hexKey, sqKey, hpelKey, qc0, qpelKey, lut);
            g_stats[ST_RESUME0 + resume]++;
        }
        /* ---- one pass at a time from the stage the fast path stopped at ---- */
        if (resume != LA_RESUME_DONE)
        {
            bool more = resume == LA_RESUME_HEX3;
            if (resume == LA_RESUME_HEX6)
            {
                for (int q = 0; q < 8; q++)
                {
                    valid[q] = q < 6;
                    if (!valid[q]) continue;
                    int qx = (s.bmx + la_hex2x(q + 1)) * 4, qy = (s.bmy + la_hex2y(q + 1)) * 4;
                    cost[q] = evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(c->mvcost, s, qx, qy);
                }
                more = la_upd_hex6(s, Pass::minKey(cost, valid, 0, 8));
            }
            int guard = 0;
            while (more && guard++ < 16)
            {
                for (int q = 0; q < 8; q++)
                {
                    valid[q] = q < 3;
                    if (!valid[q]) continue;
                    int qx = (s.bmx + la_hex2x(s.dir + q)) * 4, qy = (s.bmy + la_hex2y(s.dir + q)) * 4;
                    cost[q] = evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(c->mvcost, s, qx, qy);
                }
                more = la_upd_hex3(s, Pass::minKey(cost, valid, 0, 8));
            }
            bool subpel = true;
            if (resume <= LA_RESUME_SQ8)
            {
                for (int q = 0; q < 8; q++)
                {
                    valid[q] = true;
                    int qx = (s.bmx + la_sq1x(q + 1)) * 4, qy = (s.bmy + la_sq1y(q + 1)) * 4;
                    cost[q] = evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(c->mvcost, s, qx, qy);
                }
                subpel = la_upd_sq8(s, Pass::minKey(cost, valid, 0, 8), c->mvcost);
            }
            if (subpel)
            {
                if (resume <= LA_RESUME_HPEL)
                {
                    for (int q = 0; q < 8; q++)
                    {
                        valid[q] = q < 4;
                        if (!valid[q]) continue;
                        int qx = s.bmx + la_sq1x(q + 1) * 2, qy = s.bmy + la_sq1y(q + 1) * 2;
                        cost[q] = evalQ(fb, ref, off, qx, qy, 0) + la_mvcost(c->mvcost, s, qx, qy);
                    }
                    la_upd_hpel(s, Pass::minKey(cost, valid, 0, 8));
                }
                for (int q = 0; q < 8; q++)
                {
                    valid[q] = q >= 1 && q < 5;
                    if (q >= 5) continue;
                    int qx = s.bmx + la_sq1x(q), qy = s.bmy + la_sq1y(q);
                    cost[q] = evalQ(fb, ref, off, qx, qy, 1) + la_mvcost(c->mvcost, s, qx, qy);
                }
                la_upd_qpel(s, cost[0], Pass::minKey(cost, valid, 0, 8));
            }
        }
        { int oc = s.outcost; la_finish_skip(s); if (oc != s.outcost) g_stats[ST_SKIP]++; }
        if (s.outx == s.mvpx && s.outy == s.mvpy) g_stats[ST_OUT_EQ_MVP]++;
        listCost[i] = s.outcost;
        fenc->mvCosts[i][dist][cuXY] = s.outcost;
        mv[cuXY].x = (int16_t)s.outx; mv[cuXY].y = (int16_t)s.outy;
    }

    int bi0 = LA_COST_MAX, bi1 = LA_COST_MAX;
    if (bidir)
    {
        pixel b0[64], b1[64], avg[64];
        ola_mv m0 = fenc->mvs[0][d0 - 1][cuXY], m1 = fenc->mvs[1][d1 - 1][cuXY];
        ola_lowres_mc(r0.plane, g.stride, off, m0.x, m0.y, b0);
        ola_lowres_mc(r1.plane, g.stride, off, m1.x, m1.y, b1);
        ola_pixelavg8x8(avg, 8, b0, 8, b1, 8);
        bi0 = ola_satd8x8(fb, 8, avg, 8);
        ola_pixelavg8x8(avg, 8, r0.plane[0] + off, g.stride, r1.plane[0] + off, g.stride);
        bi1 = ola_satd8x8(fb, 8, avg, 8);
    }
    LaCuResult res = la_cu_finish(cuX, cuY, W, H, bidir, listCost[0], listCost[1], bi0, bi1, fenc->intraCost[cuXY],
                                  fenc->invQscale != NULL, fenc->invQscale ? fenc->invQscale[cuXY] : 256);
    if (res.scored)
    {
        acc.costEst += res.bcost;
        acc.costEstAq += res.bcostAq;
        acc.intraMbs += res.intraMb;
    }
    fenc->rowSatds[d0][d1][cuY] += res.bcostAq;
    fenc->lowresCosts[d0][d1][cuXY] = res.lowresCost;
}

} // namespace

extern "C" long long* emul_stats(void) { return g_stats; }

extern "C" int emul_check_tables(void)
{
    static const int hex2[8][2] = { { -1, -2 }, { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 }, { -2, 0 } };
    static const int mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };
    static const int square1[9][2] = { { 0, 0 }, { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 }, { -1, -1 }, { -1, 1 }, { 1, -1 }, { 1, 1 } };
    int bad = 0;
    for (int i = 0; i < 8; i++)
        bad += la_hex2x(i) != hex2[i][0] || la_hex2y(i) != hex2[i][1] || la_mod6m1(i) != mod6m1[i];
    for (int i = 0; i < 9; i++)
        bad += la_sq1x(i) != square1[i][0] || la_sq1y(i) != square1[i][1];
    for (int x = -300; x <= 300; x += 7)
        for (int y = -300; y <= 300; y += 11)
        {
            int p = la_pack_mv(x, y);
            bad += la_mv_x(p) != x || la_mv_y(p) != y;
        }
    return bad;
}

extern "C" int64_t emul_estimate(ola_ctx* c, ola_frame* fenc, ola_frame* ref0, ola_frame* ref1, int d0, int d1,
                                 int search0, int search1, int sliced, int weightp, const ola_weight* weight, ola_weight* usedWeight)
{
    const ola_geom& g = fenc->g;
    int doSearch[2];
    doSearch[0] = search0 >= 0 ? search0 : (d0 > 0 && fenc->mvs[0][d0 - 1][0].x == OLA_MV_SENTINEL);
    doSearch[1] = search1 >= 0 ? search1 : (d1 > 0 && fenc->mvs[1][d1 - 1][0].x == OLA_MV_SENTINEL);
    ola_weight w = { 0, 0, 0, 0 };
    if (weight)
    {
        w = *weight;
        if (w.present) ola_apply_weight(c, ref0, &w);
    }
    else if (weightp && doSearch[0])
        ola_weights_analyse(c, fenc, ref0, &w);
    if (usedWeight) *usedWeight = w;
    Refs wref0;
    wref0.stride = g.stride;
    for (int i = 0; i < 4; i++)
        wref0.plane[i] = w.present ? c->wbuffer[i] + g.padOffset : ref0->plane[i];

    Acc total = { 0, 0, 0 };
    int useSlices = sliced && c->numCoopSlices > 1 && (d1 > 0 || doSearch[0] || doSearch[1]);
    int nSlices = useSlices ? c->numCoopSlices : 1;
    for (int sl = 0; sl < nSlices; sl++)
    {
        int firstY = useSlices ? c->numRowsPerSlice * sl : 0;
        int lastY = (!useSlices || sl == nSlices - 1) ? g.hCU - 1 : c->numRowsPerSlice * (sl + 1) - 1;
        int lastRow = 1;
        for (int cuY = lastY; cuY >= firstY; cuY--)
        {
            fenc->rowSatds[d0][d1][cuY] = 0;
            for (int cuX = g.wCU - 1; cuX >= 0; cuX--)
                emulCU(c, fenc, ref0, ref1, wref0, cuX, cuY, d0, d1, doSearch, lastRow, total);
            lastRow = 0;
        }
    }
    fenc->costEstAq[d0][d1] = total.costEstAq;
    if (d1 == 0)
        fenc->intraMbs[d0] += total.intraMbs;
    int64_t score = total.costEst;
    if (d1 > 0)
        score = score * 100 / (130 + c->bFrameBias);
    fenc->costEst[d0][d1] = score;
    return score;
}
