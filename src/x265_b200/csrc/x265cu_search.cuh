/* x265cu_search.cuh -- the motion-search kernels of a frame-cost estimate (sm_100a).
 *
 * Replace, for ONE reference list of one estimate, the search half of
 * CostEstimateGroup::estimateCUCost (encoder/slicetype.cpp:2106-2160) and the lowres branch of
 * MotionEstimate::motionEstimate (encoder/motion.cpp:571-1172).  The L0 and L1 searches of an
 * estimate are independent of each other (they only meet in the bidir/intra decision, which
 * cost_kernel does afterwards for the whole frame in parallel), so the unit of work is a SEARCH =
 * (job, list).
 *
 * The reference's dependency: a CU's MVP candidates are the MVs of its right, below, below-left and
 * below-right neighbours (slicetype.cpp:2117-2128), which makes a frame (or cooperative slice) a
 * wavefront and every CU row a serial chain of ~10 dependent measuring passes per CU.  Here the
 * chain only carries DECISIONS; the measuring is taken off it, bit-exactly, in three steps:
 *
 * 1. ONE-SHOT SEARCH (search_mv).  Everything the search does once the MVP is chosen is a function
 *    of the MVP alone, and in the common case (the search confirms its predictor) it only visits
 *    positions that are known up front.  All 26 of them are measured in one straight-line burst
 *    from a window staged in shared memory, and la_fast_path() (la_core.h) replays the reference's
 *    decisions on the costs; only when the search really moves does it fall back to
 *    one-pass-at-a-time evaluation, from the stage where the speculation stopped applying.
 *
 * 2. REFINE KERNEL (refine_kernel), run a few times: every CU of every search in parallel, no
 *    dependencies inside a launch.  Iteration k reads an ESTIMATE H_k of the search's MV field (H_0 = a
 *    hint: the field an earlier, temporally adjacent search produced; zero when there is none), takes a
 *    CU's neighbour MVs from it, runs the CU's full search for every such vector it has not measured
 *    yet and memoises it per CU as {MVP, SATD at the MVP, resulting MV, resulting cost}; then it
 *    evaluates the CU exactly as the reference would IF its neighbours had those MVs, which gives
 *    H_k+1.  New vectors that this frame's pixels produce thus reach their neighbours' memos one CU
 *    further per iteration -- in parallel instead of on the chain.  Estimates only ever select WHICH
 *    vectors get measured early, never a result.
 *
 * 3. COMMIT KERNEL (search_kernel): the wavefront.  One warp per CU row; per CU it takes the REAL
 *    neighbour MVs, looks their SATDs up in the memo (same MV => same pixels => same cost), picks
 *    the MVP exactly as the reference does (strict <, reference order, skipCost rule) and takes the
 *    memoised search result.  Pixels are only touched on the chain when a vector was never
 *    predicted (then: the reference's CAND pass and/or a one-shot search, inline).
 *
 * A finished CU publishes ONE 64-bit word {tag = 1, packed MV}; tag and data travel in the same
 * naturally aligned 8-byte store, so there is no separate progress counter and no fence on the
 * chain.  Inside a row group the words live in shared memory, between groups in a global hand-off
 * row (L2).  A group only ever waits for a group with a LOWER block index (launched earlier), so
 * the scheme cannot deadlock even when a launch does not fit the GPU.
 *
 * Lane mapping of every measure: quad q = lane >> 2 owns one candidate block, each lane its 4x4
 * sub-block; winners are warp min-reductions over packed (cost << 3 | k) keys (la_core.h).
 */
#ifndef X265CU_SEARCH_CUH
#define X265CU_SEARCH_CUH

/* one (job, list) search of a batch */
struct SearchPlan
{
    int job, list;
    int rowsPerSlice, numSlices;   /* cooperative slices (numSlices == 1: whole frame) */
    const int* hint;               /* packed MV field predicting this search's result, or NULL (zero field) */
    int hintNeg;                   /* the hint is the opposite list's field: negate it */
    int4* memo;                    /* [nCU][MEMO_N] {mvp, SATD at mvp or -1, result MV, result cost or -1 (empty)} */
    int* field[3];                 /* [nCU] estimates of the result field: refine iteration k writes field[k % 3] */
};

/* one row group of a search: a CTA of the commit kernel */
struct SearchItem
{
    int search;                    /* index into SearchPlan[] */
    int sliceFirstY, sliceLastY;   /* cooperative slice (or whole frame) this group belongs to */
    int firstY, lastY;             /* CU rows of this group (lastY = bottom row, processed first) */
    int pubBase;                   /* hand-off row this group's TOP row publishes to (index of entry 0), -1: none */
    int subBase;                   /* hand-off row the group's BOTTOM row reads (the group below), -1: none */
};

#define MEMO_N 8
#define SEARCH_MAX_GROUP_ROWS 16
#define SEARCH_MAX_THREADS 256
#define SPEC_WARPS 8
#define HAND_TAG (1ull << 32)

/* control words behind the hand-off rows (zeroed with them before every launch).  A CTA takes its work item from a
 * ticket instead of blockIdx: items are ordered so that an item only waits for items before it, and an item is only
 * ever taken by a CTA that is already running, so every wait ends whatever order the hardware starts blocks in. */
struct SearchCtl { unsigned int ticket[4]; unsigned int error; unsigned int pad[3]; };

__device__ __forceinline__ int take_ticket(unsigned int* counter)
{
    __shared__ int sTicket;
    if (threadIdx.x == 0) sTicket = (int)atomicAdd(counter, 1u);
    __syncthreads();
    return sTicket;
}

/* optional instrumentation for kernel tuning (compile with -DX265CU_SEARCH_STATS; never in the shipped library) */
#ifdef X265CU_SEARCH_STATS
__device__ unsigned long long g_searchStats[32];
#define SSTAT_ADD(i, v) do { if ((threadIdx.x & 31) == 0) atomicAdd(&g_searchStats[i], (unsigned long long)(v)); } while (0)
#define SSTAT_CLOCK() clock64()
#define SSTAT_T(v) const long long v = clock64()
/* per-row step timeline of the commit kernel: [cuY][event] = {globaltimer ns, x0 | kind << 16 | n << 24} */
__device__ unsigned long long g_traceT[256][512];
__device__ unsigned int g_traceE[256][512];
__device__ int g_traceN[256];
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define STRACE(row, x0, kind, n) do { if ((threadIdx.x & 31) == 0) { int i_ = g_traceN[row]; if (i_ < 512) { g_traceT[row][i_] = gtimer(); g_traceE[row][i_] = (unsigned)(x0) | ((unsigned)(kind) << 16) | ((unsigned)(n) << 24); g_traceN[row] = i_ + 1; } } } while (0)
#else
#define STRACE(row, x0, kind, n) do { } while (0)
#define SSTAT_ADD(i, v) do { } while (0)
#define SSTAT_CLOCK() 0ll
#define SSTAT_T(v) do { } while (0)
#endif

/* two-source fetch at quarter-pel MV (qx, qy); when the MV is not odd both sources coincide and the
 * rounded average returns the sample itself, so the code is the same for every candidate
 * (ReferencePlanes::lowresMC / lowresQPelCost, common/lowres.h:62-103) */
template <typename P>
__device__ __forceinline__ void fetch_qpel(const P* __restrict__ refLane, int planeSize, int stride, int qx, int qy, typename Px<P>::Row4 out[4])
{
    const int hpelA = (qy & 2) | ((qx & 2) >> 1);
    const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
    const int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
    const P* a = refLane + (hpelA * planeSize + (qy >> 2) * stride + (qx >> 2));
    const P* b = refLane + (hpelB * planeSize + (qy2 >> 2) * stride + (qx2 >> 2));
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = Px<P>::avg(Px<P>::load(a + i * stride), Px<P>::load(b + i * stride));
}

/* full-pel position at sample offset `off` from the lane's block: plane 0 only */
template <typename P>
__device__ __forceinline__ void fetch_off(const P* __restrict__ refLane, int stride, int off, typename Px<P>::Row4 out[4])
{
    const P* a = refLane + off;
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = Px<P>::load(a + i * stride);
}

__device__ __forceinline__ uint32_t warp_min_key(bool valid, int cost, int q)
{
    uint32_t key = valid ? la_key(cost, q) : LA_KEY_NONE;
    return __reduce_min_sync(FULL_MASK, key);
}

/* wait for a published hand-off word and return its MV (plain spinning: backing off with nanosleep was measured
 * and changes nothing -- the schedulers already favour the warps that have work) */
__device__ __forceinline__ int hand_wait(volatile const unsigned long long* e)
{
    unsigned long long v;
    do { v = *e; } while (!(v & HAND_TAG));
    return (int)(uint32_t)v;
}

/* per-lane candidate geometry of the fixed-shape rounds/passes (motion.cpp:64-66 tables), computed once */
struct LaneGeom
{
    int q;
    /* one-shot rounds */
    int r1dx, r1dy;          /* round 1: quads 3..6 = half-pel square around pm (quarter-pel units) */
    int r2off, r2dx, r2dy;   /* round 2: quads 0..5 = hexagon, 6..7 = square1[1..2] */
    int r3off, r3dx, r3dy;   /* round 3: quads 0..5 = square1[3..8] */
    int r4dx, r4dy;          /* round 4: quads 0..4 = pm + square1[0..4] (quarter-pel units) */
    /* one-pass-at-a-time fallback */
    int hex6off, hex6dx, hex6dy, sq8off, sq8dx, sq8dy, hpdx, hpdy, qpdx, qpdy;
};

__device__ __forceinline__ LaneGeom lane_geom(int lane, int stride)
{
    LaneGeom L;
    const int q = lane >> 2;
    L.q = q;
    L.r1dx = (q >= 3 && q < 7) ? la_sq1x(q - 2) * 2 : 0;
    L.r1dy = (q >= 3 && q < 7) ? la_sq1y(q - 2) * 2 : 0;
    L.r2dx = q < 6 ? la_hex2x(q + 1) : la_sq1x(q - 5);
    L.r2dy = q < 6 ? la_hex2y(q + 1) : la_sq1y(q - 5);
    L.r2off = L.r2dy * stride + L.r2dx;
    L.r3dx = q < 6 ? la_sq1x(q + 3) : 0;
    L.r3dy = q < 6 ? la_sq1y(q + 3) : 0;
    L.r3off = L.r3dy * stride + L.r3dx;
    L.r4dx = q < 5 ? la_sq1x(q) : 0;
    L.r4dy = q < 5 ? la_sq1y(q) : 0;
    L.hex6dx = la_hex2x((q + 1) & 7); L.hex6dy = la_hex2y((q + 1) & 7);
    L.hex6off = L.hex6dy * stride + L.hex6dx;
    L.sq8dx = la_sq1x(q + 1); L.sq8dy = la_sq1y(q + 1);
    L.sq8off = L.sq8dy * stride + L.sq8dx;
    L.hpdx = la_sq1x((q + 1) & 7) * 2; L.hpdy = la_sq1y((q + 1) & 7) * 2;
    L.qpdx = la_sq1x(q); L.qpdy = la_sq1y(q);
    return L;
}

/* The window a one-shot search reads: WIN_H rows x WIN_W samples of each of the four hpel planes
 * around the clipped MVP pm, staged in shared memory (one private buffer per warp) with coalesced
 * row loads.  Origin (full-pel, relative to the CU): wx0 = ((pm.x >> 2) - 2) & ~3, wy0 = (pm.y >> 2) - 2.
 * Covers the hexagon (+-2) and the square (+-1) around round(pm) and every half-/quarter-pel point
 * within +-2 quarter samples of pm, i.e. all 26 positions of the no-move path. */
#ifndef WIN_W
#define WIN_W 16
#define WIN_H 13
#endif
#define WIN_ROW_UNITS (WIN_W / 4)
#define WIN_PLANE_UNITS (WIN_H * WIN_ROW_UNITS)
#define WIN_UNITS (4 * WIN_PLANE_UNITS)
#define WIN_PITCH (WIN_UNITS + 8)           /* + slack: an aligned read's second unit may lie one past the end */

/* 4 samples at window position (dx, dy) of plane `plane`; RU = units of 4 samples per window row (the TMA variant of the
 * plain kernel loads wider rows: its boxes must start on 16-byte boundaries) */
template <typename P, int RU = WIN_ROW_UNITS>
__device__ __forceinline__ typename Px<P>::Row4 win_read(const typename Px<P>::Row4* win, int plane, int dx, int dy)
{
    const typename Px<P>::Row4* u = win + plane * (WIN_H * RU) + dy * RU + (dx >> 2);
    return Px<P>::combine(u[0], u[1], dx & 3);
}

/* this lane's 4x4 of the reference block at quarter-pel MV (qx, qy), from the window (lx, ly = the
 * lane's sub-block offset minus the window origin) */
template <typename P, int RU = WIN_ROW_UNITS>
__device__ __forceinline__ void win_qpel(const typename Px<P>::Row4* win, int lx, int ly, int qx, int qy, typename Px<P>::Row4 out[4])
{
    const int hpelA = (qy & 2) | ((qx & 2) >> 1);
    const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
    const int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
    const int ax = lx + (qx >> 2), ay = ly + (qy >> 2), bx = lx + (qx2 >> 2), by = ly + (qy2 >> 2);
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = Px<P>::avg(win_read<P, RU>(win, hpelA, ax, ay + i), win_read<P, RU>(win, hpelB, bx, by + i));
}

template <typename P, int RU = WIN_ROW_UNITS>
__device__ __forceinline__ void win_fpel(const typename Px<P>::Row4* win, int lx, int ly, int fx, int fy, typename Px<P>::Row4 out[4])
{
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = win_read<P, RU>(win, 0, lx + fx, ly + fy + i);
}

/* The whole search of one CU for a given MVP `m` (packed quarter-pel), result before the skip
 * shortcut: motion.cpp:587-624,670-742,1081-1119.  Also returns the SATD at pm without mvcost
 * (centerSatd): when pm == m (candOk) that is exactly the candidate cost of the neighbour MV m in
 * the MVP selection (slicetype.cpp:2130-2150), so the CAND pass costs nothing extra. */
template <typename P>
__device__ __forceinline__ void search_mv(const P* __restrict__ refCU, const P* __restrict__ refLane, typename Px<P>::Row4* win,
                                          int planeSize, int stride, const typename Px<P>::Row4 fe[4],
                                          const uint16_t* __restrict__ lut, const LaneGeom& L, int lane, int bx, int by,
                                          int cuX, int cuY, int W, int H, int m,
                                          int& outMv, int& outCost, int& centerSatd, bool& candOk)
{
    typedef typename Px<P>::Row4 Row4;
    const int q = L.q;
    LaSearch s;
    la_search_begin(s, cuX, cuY, W, H, 0, 0, 0, 0, 0, 0);
    s.mvpx = la_mv_x(m); s.mvpy = la_mv_y(m);
    la_enter_start(s);
    candOk = s.pmx == s.mvpx && s.pmy == s.mvpy;
    const uint16_t* __restrict__ lutx = lut - s.mvpx;
    const uint16_t* __restrict__ luty = lut - s.mvpy;
    const int bm0x = (s.pmx + 2) >> 2, bm0y = (s.pmy + 2) >> 2;
    const int wx0 = ((s.pmx >> 2) - 2) & ~3, wy0 = (s.pmy >> 2) - 2;

    /* ---- stage the window: unit i = (plane, row, 4-sample column), consecutive lanes walk a row ---- */
    __syncwarp();
    {
        const P* __restrict__ wbase = refCU + wy0 * stride + wx0;
#pragma unroll
        for (int k = 0; k < (WIN_UNITS + 31) / 32; k++)
        {
            const int i = lane + 32 * k;
            if (i < WIN_UNITS)
            {
                const int plane = i / WIN_PLANE_UNITS, rem = i - plane * WIN_PLANE_UNITS;
                win[i] = Px<P>::load_aligned(wbase + plane * planeSize + (rem / WIN_ROW_UNITS) * stride + (rem % WIN_ROW_UNITS) * 4);
            }
        }
    }
    __syncwarp();
    const int lx = bx - wx0, ly = by - wy0;

    /* ---- round 1: q0 qpel MVP (no mvcost), q1 rounded MVP, q3..6 the 4 half-pel points around pm ---- */
    int cost1;
    {
        const int qx = (q == 1 ? bm0x * 4 : s.pmx) + L.r1dx;
        const int qy = (q == 1 ? bm0y * 4 : s.pmy) + L.r1dy;
        Row4 r[4];
        win_qpel<P>(win, lx, ly, qx, qy, r);
        const int mvc = q == 0 ? 0 : lutx[qx] + luty[qy];
        cost1 = quad_sum(sad4x4<P>(fe, r)) + mvc;
    }
    /* ---- rounds 2, 3: hexagon (6) and unit square (8) around the rounded MVP, full-pel SAD;
     *      the two spare quads of round 3 measure the zero MV straight from global memory ---- */
    int cost2, cost3;
    {
        Row4 r2[4], r3[4];
        win_fpel<P>(win, lx, ly, bm0x + L.r2dx, bm0y + L.r2dy, r2);
        if (q < 6) win_fpel<P>(win, lx, ly, bm0x + L.r3dx, bm0y + L.r3dy, r3);
        else fetch_off<P>(refLane, stride, 0, r3);
        const int f3x = q < 6 ? bm0x + L.r3dx : 0, f3y = q < 6 ? bm0y + L.r3dy : 0;
        cost2 = quad_sum(sad4x4<P>(fe, r2)) + lutx[(bm0x + L.r2dx) * 4] + luty[(bm0y + L.r2dy) * 4];
        cost3 = quad_sum(sad4x4<P>(fe, r3)) + lutx[f3x * 4] + luty[f3y * 4];
    }
    /* ---- round 4: SATD at pm and the 4 quarter-pel points around it ---- */
    int cost4, raw4;
    {
        const int qx = s.pmx + L.r4dx, qy = s.pmy + L.r4dy;
        Row4 r[4];
        win_qpel<P>(win, lx, ly, qx, qy, r);
        raw4 = quad_sum(satd4x4_abs<P>(fe, r)) >> 1;
        cost4 = raw4 + lutx[qx] + luty[qy];
    }
    int resume;
    {
        const int c0 = __shfl_sync(FULL_MASK, cost1, 0), c1 = __shfl_sync(FULL_MASK, cost1, 4), c2 = __shfl_sync(FULL_MASK, cost3, 24);
        const uint32_t hpelKey = warp_min_key(q >= 3 && q < 7, cost1, q - 3);
        const uint32_t hexKey = warp_min_key(q < 6, cost2, q);
        const uint32_t sqKey = __reduce_min_sync(FULL_MASK, q < 6 ? la_key(cost3, q + 2) : la_key(cost2, q - 6));
        const int qc0 = __shfl_sync(FULL_MASK, cost4, 0);
        centerSatd = __shfl_sync(FULL_MASK, raw4, 0);
        /* (only for a vector that is a plausible neighbour MV, i.e. at most a CU + search slack beyond the clipped one:
         * estimates are not trusted to point inside the padded planes) */
        if (!candOk && abs(s.mvpx - s.pmx) <= 96 && abs(s.mvpy - s.pmy) <= 96)
        {
            /* the MVP was clipped (picture edge): the neighbour's own vector is measured where it points, unclipped
             * (slicetype.cpp:2138), straight from global memory */
            Row4 rr[4];
            fetch_qpel<P>(refLane, planeSize, stride, s.mvpx, s.mvpy, rr);
            centerSatd = __shfl_sync(FULL_MASK, quad_sum(satd4x4_abs<P>(fe, rr)) >> 1, 0);
            candOk = true;
        }
        const uint32_t qpelKey = warp_min_key(q >= 1 && q < 5, cost4, q);
        resume = la_fast_path(s, c0, c1, c2, hexKey, sqKey, hpelKey, qc0, qpelKey, lut);
        SSTAT_ADD(4 + resume, 1);
    }

    /* ---- the search moved: continue one pass at a time from the stage the speculation stopped at ---- */
    if (resume != LA_RESUME_DONE)
    {
        bool more = resume == LA_RESUME_HEX3;
        if (resume == LA_RESUME_HEX6)
        {
            const int fx = s.bmx + L.hex6dx, fy = s.bmy + L.hex6dy;
            Row4 r[4];
            fetch_off<P>(refLane, stride, s.bmy * stride + s.bmx + L.hex6off, r);
            const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
            more = la_upd_hex6(s, warp_min_key(q < 6, cost, q));
        }
        while (more)
        {
            const int hdx = la_hex2x((s.dir + q) & 7), hdy = la_hex2y((s.dir + q) & 7);
            const int hx = s.bmx + hdx, hy = s.bmy + hdy;
            Row4 r3[4];
            fetch_off<P>(refLane, stride, hy * stride + hx, r3);
            const int c3 = quad_sum(sad4x4<P>(fe, r3)) + lutx[hx * 4] + luty[hy * 4];
            more = la_upd_hex3(s, warp_min_key(q < 3, c3, q));
        }
        bool subpel = true;
        if (resume <= LA_RESUME_SQ8)
        {
            const int fx = s.bmx + L.sq8dx, fy = s.bmy + L.sq8dy;
            Row4 r[4];
            fetch_off<P>(refLane, stride, s.bmy * stride + s.bmx + L.sq8off, r);
            const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
            subpel = la_upd_sq8(s, warp_min_key(true, cost, q), lut);
        }
        if (subpel)
        {
            if (resume <= LA_RESUME_HPEL)
            {
                const int qx = s.bmx + L.hpdx, qy = s.bmy + L.hpdy;
                Row4 r[4];
                fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[qx] + luty[qy];
                la_upd_hpel(s, warp_min_key(q < 4, cost, q));
            }
            {
                const int qx = s.bmx + L.qpdx, qy = s.bmy + L.qpdy;
                Row4 r[4];
                fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                const int cost = (quad_sum(satd4x4_abs<P>(fe, r)) >> 1) + lutx[qx] + luty[qy];
                const int c0 = __shfl_sync(FULL_MASK, cost, 0);
                la_upd_qpel(s, c0, warp_min_key(q >= 1 && q < 5, cost, q));
            }
        }
    }
    outMv = la_pack_mv(s.outx, s.outy);
    outCost = s.outcost;
}

/* ===========================================================================================
 * refine_kernel: grid (ceil(nCU / SPEC_WARPS), searches), one warp per CU, no dependencies.
 * Lane k (and k + 8, k + 16, k + 24) holds memo entry k of the CU, so a lookup is one ballot.
 * =========================================================================================== */
__device__ __forceinline__ int hint_at(const int* __restrict__ hint, int neg, int idx)
{
    if (!hint) return 0;
    const int h = __ldg(hint + idx);
    return neg ? la_pack_mv(-la_mv_x(h), -la_mv_y(h)) : h;
}

/* index of the memo entry for vector v (-1: none); e = this lane's entry */
__device__ __forceinline__ int memo_find(const int4& e, int v)
{
    return __ffs(__ballot_sync(FULL_MASK, e.w >= 0 && e.x == v) & ((1u << MEMO_N) - 1)) - 1;
}

template <typename P>
__global__ void __launch_bounds__(SPEC_WARPS * 32, 3)
refine_kernel(const JobDev* __restrict__ jobs, const SearchPlan* __restrict__ plans, GeomDev g, const uint16_t* __restrict__ lut, int iter)
{
    __shared__ typename Px<P>::Row4 sWin[SPEC_WARPS][WIN_PITCH];
    const SearchPlan pl = plans[blockIdx.y];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = g.wCU, H = g.hCU;
    const int cuXY = blockIdx.x * SPEC_WARPS + warp;
    if (cuXY >= g.nCU) return;
    const int cuY = cuXY / W, cuX = cuXY - cuY * W;
    const JobDev* __restrict__ jp = jobs + pl.job;
    const int bidir = jp->bidir;

    /* the current estimate of the neighbours' MVs, in the reference's candidate order (slicetype.cpp:2117-2128);
     * bottom row of a cooperative slice (or of the frame): no candidates from below (slicetype.cpp:1957-1968) */
    const bool lastRow = cuY == H - 1 || (pl.numSlices > 1 && (cuY + 1) % pl.rowsPerSlice == 0 && (cuY + 1) / pl.rowsPerSlice < pl.numSlices);
    const int* __restrict__ fin = iter == 0 ? pl.hint : pl.field[(iter - 1) % 3];
    const int neg = iter == 0 ? pl.hintNeg : 0;
    int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0, numc = 0;
    if (cuX < W - 1) { nb0 = hint_at(fin, neg, cuXY + 1); numc = 1; }
    if (!lastRow)
    {
        const int mb = hint_at(fin, neg, cuXY + W);
        if (numc == 0) nb0 = mb; else nb1 = mb;
        numc++;
        if (cuX > 0) { const int bl = hint_at(fin, neg, cuXY + W - 1); if (numc == 1) nb1 = bl; else nb2 = bl; numc++; }
        if (cuX < W - 1) { const int br = hint_at(fin, neg, cuXY + W + 1); if (numc == 2) nb2 = br; else nb3 = br; numc++; }
    }

    if (iter > 0)
    {
        /* same neighbour estimates as in the previous iteration: same memo, same outcome -- nothing to do */
        const int* __restrict__ fprev = iter == 1 ? pl.hint : pl.field[(iter - 2) % 3];
        const int nprev = iter == 1 ? pl.hintNeg : 0;
        int p0 = 0, p1 = 0, p2 = 0, p3 = 0, pc = 0;
        if (cuX < W - 1) { p0 = hint_at(fprev, nprev, cuXY + 1); pc = 1; }
        if (!lastRow)
        {
            const int mb = hint_at(fprev, nprev, cuXY + W);
            if (pc == 0) p0 = mb; else p1 = mb;
            pc++;
            if (cuX > 0) { const int bl = hint_at(fprev, nprev, cuXY + W - 1); if (pc == 1) p1 = bl; else p2 = bl; pc++; }
            if (cuX < W - 1) { const int br = hint_at(fprev, nprev, cuXY + W + 1); if (pc == 2) p2 = br; else p3 = br; pc++; }
        }
        if (p0 == nb0 && p1 == nb1 && p2 == nb2 && p3 == nb3)
        {
            if (lane == 0) pl.field[iter % 3][cuXY] = fin[cuXY];
            return;
        }
    }
    int4* __restrict__ memo = pl.memo + (size_t)cuXY * MEMO_N;
    int4 e = make_int4(0, -1, 0, -1);
    if (iter > 0) e = memo[lane & (MEMO_N - 1)];
    int n = __popc(__ballot_sync(FULL_MASK, e.w >= 0) & ((1u << MEMO_N) - 1));

    /* ---- measure every neighbour vector the memo does not hold yet (no candidates at all: the MVP is zero) ---- */
    const int nWant = numc ? numc : 1;
    bool loaded = false;
    typename Px<P>::Row4 fe[4];
    const P* __restrict__ refPlane = (const P*)(pl.list ? jp->ref1 : jp->ref0w);
    const int sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int stride = g.stride, planeSize = (int)g.planeSize;
    const int rowBase = (8 * cuY + by) * stride + bx;
    const LaneGeom L = lane_geom(lane, stride);
#pragma unroll 1
    for (int i = 0; i < nWant; i++)
    {
        const int v = i == 0 ? nb0 : (i == 1 ? nb1 : (i == 2 ? nb2 : nb3));
        if (memo_find(e, v) >= 0 || n >= MEMO_N) continue;
        if (!loaded)
        {
            const P* __restrict__ fencPlane = (const P*)jp->fenc;
#pragma unroll
            for (int y = 0; y < 4; y++)
                fe[y] = Px<P>::load_aligned(fencPlane + rowBase + 8 * cuX + y * stride);
            loaded = true;
        }
        int oMv, oCost, cs;
        bool cok;
        search_mv<P>(refPlane + 8 * cuY * stride + 8 * cuX, refPlane + rowBase + 8 * cuX, sWin[warp], planeSize, stride, fe, lut, L, lane, bx, by,
                     cuX, cuY, W, H, v, oMv, oCost, cs, cok);
        const int4 ne = make_int4(v, cok ? cs : -1, oMv, oCost);
        if ((lane & (MEMO_N - 1)) == n) e = ne;
        if (lane == n) memo[n] = ne;
        n++;
    }
    if (iter == 0 && lane < MEMO_N && lane >= n) memo[lane] = e;     /* empty entries */

    /* ---- what the reference would decide if the neighbours had these MVs: the next estimate ---- */
    int outMv = fin ? hint_at(fin, neg, cuXY) : 0;
    {
        const int i0 = memo_find(e, nb0), i1 = memo_find(e, nb1), i2 = memo_find(e, nb2), i3 = memo_find(e, nb3);
        const int k0 = __shfl_sync(FULL_MASK, e.y, i0 & 31), k1 = __shfl_sync(FULL_MASK, e.y, i1 & 31);
        const int k2 = __shfl_sync(FULL_MASK, e.y, i2 & 31), k3 = __shfl_sync(FULL_MASK, e.y, i3 & 31);
        const bool ok = (i0 >= 0 && (k0 >= 0 || numc == 0)) && (numc < 2 || (i1 >= 0 && k1 >= 0)) && (numc < 3 || (i2 >= 0 && k2 >= 0)) && (numc < 4 || (i3 >= 0 && k3 >= 0));
        if (ok)
        {
            LaSearch s;
            la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);
            la_upd_cand(s, k0, k1, k2, k3);
            const int im = memo_find(e, la_pack_mv(s.mvpx, s.mvpy));     /* the MVP is one of the candidates (or zero): always found */
            const int rMv = __shfl_sync(FULL_MASK, e.z, im & 31), rCost = __shfl_sync(FULL_MASK, e.w, im & 31);
            if (im >= 0)
            {
                s.outx = la_mv_x(rMv); s.outy = la_mv_y(rMv); s.outcost = rCost;
                la_finish_skip(s);
                outMv = la_pack_mv(s.outx, s.outy);
            }
        }
    }
    if (lane == 0) pl.field[iter % 3][cuXY] = outMv;
}

/* ===========================================================================================
 * search_kernel (commit): one CTA per row group, one warp per CU row, wavefront over the rows.
 *
 * A row is a serial chain (each CU needs its right neighbour's MV), but the refined estimate
 * predicts that MV almost everywhere, so the warp commits up to 32 CUs per STEP: lane i takes CU
 * x0 - i, assumes its right neighbour's MV is the estimate (lane 0 knows the true one), reads the
 * true MVs of the row below, and evaluates the CU from its memo entirely on its own (candidate
 * SATDs, MVP by the reference's strict-< chain, memoised search result, skip rule).  Then the
 * assumptions are checked against the actual results of the lanes to the right: the longest prefix
 * of lanes whose inputs were all true is final and is published at once; the chain restarts at
 * the first lane that assumed wrongly, waited for the row below, or met a vector its memo does
 * not hold (that CU is then handled by the whole warp: CAND pass / one-shot search).
 * =========================================================================================== */
template <typename P>
__host__ __device__ inline size_t search_smem_bytes(int rows, int wCU)
{
    return (size_t)rows * wCU * (sizeof(unsigned long long) + sizeof(int)) + (size_t)rows * WIN_PITCH * sizeof(typename Px<P>::Row4);
}

template <typename P>
__global__ void __launch_bounds__(SEARCH_MAX_GROUP_ROWS * 32)
search_kernel(const JobDev* __restrict__ jobs, const SearchPlan* __restrict__ plans, const SearchItem* __restrict__ items, GeomDev g,
              const uint16_t* __restrict__ lut, unsigned long long* gHand, int rowsMax, int estIdx, unsigned int* ticket)
{
    extern __shared__ unsigned long long sHand[];  /* [rowsMax][W] hand-off words, [rowsMax][W] estimates, one window per warp */
    const SearchItem it = items[take_ticket(ticket)];
    const SearchPlan pl = plans[it.search];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nRows = it.lastY - it.firstY + 1;
    const int W = g.wCU, H = g.hCU;
    int* sEstAll = (int*)(sHand + rowsMax * W);
    typename Px<P>::Row4* win = (typename Px<P>::Row4*)(sEstAll + rowsMax * W) + warp * WIN_PITCH;
    for (int i = threadIdx.x; i < nRows * W; i += blockDim.x)
    {
        sHand[i] = 0;
        const int r = i / W, x = i - r * W;
        sEstAll[i] = estIdx >= 0 ? pl.field[estIdx][(it.lastY - r) * W + x] : 0;
    }
    const JobDev* __restrict__ jp = jobs + pl.job;
    const int list = pl.list;
    __syncthreads();
    if (warp >= nRows) return;

    const int sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int stride = g.stride, planeSize = (int)g.planeSize;
    const int q = lane >> 2;
    const int bidir = jp->bidir;
    const P* __restrict__ fencPlane = (const P*)jp->fenc;
    const P* __restrict__ refPlane = (const P*)(list ? jp->ref1 : jp->ref0w);
    int* __restrict__ mvMirror = jp->mvs[list];
    int* __restrict__ mcMirror = jp->mvCosts[list];
    int* __restrict__ mvOut = jp->outMvs[list];
    int* __restrict__ mcOut = jp->outMvCosts[list];

    /* warp r owns row cuY = lastY - r; it depends on row cuY + 1 */
    const int cuY = it.lastY - warp;
    const bool lastRow = cuY == it.sliceLastY;              /* bottom row of the slice: no candidates from below */
    const bool publishGlobal = warp == nRows - 1 && it.pubBase >= 0;
    volatile unsigned long long* myHand = sHand + warp * W;
    volatile unsigned long long* myHandG = gHand + (it.pubBase >= 0 ? it.pubBase : 0);
    volatile const unsigned long long* below = (warp == 0) ? (volatile const unsigned long long*)(gHand + (it.subBase >= 0 ? it.subBase : 0))
                                                           : (volatile const unsigned long long*)(sHand + (warp - 1) * W);
    const int* sEst = sEstAll + warp * W;
    const int4* __restrict__ memoRow = pl.memo + (size_t)cuY * W * MEMO_N;
    const int rowBase = (8 * cuY + by) * stride + bx;
    int prevMv = 0;                                /* true MV of (x0 + 1, cuY) */
    int x0 = W - 1;                                /* rightmost CU not yet final */
    int serialLeft = estIdx >= 0 ? 0 : 0x7fffffff; /* CUs to take one at a time before the next wide step (no estimate: all) */
#ifdef X265CU_SEARCH_STATS
    if (lane == 0) g_traceN[cuY] = 0;
    STRACE(cuY, x0, 0, 0);
#endif
    /* the row's memo is read once per CU: pull it into L1 ahead of the chain */
    for (int i = lane; i < W * 4; i += 32)
        asm volatile("prefetch.global.L1 [%0];" :: "l"((const char*)memoRow + (size_t)i * 32));

    while (x0 >= 0)
    {
        if (serialLeft == 0)
        {
            /* ================= wide step: lane i evaluates CU x0 - i ================= */
            const int cu = x0 - lane;
            const bool valid = cu >= 0;
            const int cx = valid ? cu : 0;
            const bool hasR = cx < W - 1, hasB = !lastRow, hasBL = hasB && cx > 0, hasBR = hasB && hasR;
            /* ---- true MVs of the row below (published right to left; every word's own tag is checked) ---- */
            bool avail = true;
            int mvB = 0, mvBL = 0, mvBR = 0;
            if (hasB)
            {
                const unsigned long long wB = below[cx], wBL = below[hasBL ? cx - 1 : cx], wBR = below[hasBR ? cx + 1 : cx];
                avail = (wB & wBL & wBR & HAND_TAG) != 0;
                mvB = (int)(uint32_t)wB; mvBL = (int)(uint32_t)wBL; mvBR = (int)(uint32_t)wBR;
            }
            if (!__shfl_sync(FULL_MASK, avail, 0)) continue;          /* the chain's head waits for the row below */
            SSTAT_ADD(16, 1);

            /* ---- candidates in the reference's order (slicetype.cpp:2117-2128); the right neighbour's MV is an assumption for lanes > 0 ---- */
            const int r = lane == 0 ? prevMv : sEst[hasR ? cx + 1 : cx];
            int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0, numc = 0;
            if (hasR) { nb0 = r; numc = 1; }
            if (hasB)
            {
                if (numc == 0) nb0 = mvB; else nb1 = mvB;
                numc++;
                if (hasBL) { if (numc == 1) nb1 = mvBL; else nb2 = mvBL; numc++; }
                if (hasBR) { if (numc == 2) nb2 = mvBR; else nb3 = mvBR; numc++; }
            }
            /* ---- this lane's CU from its memo: SATD and search result of every candidate vector ---- */
            int k0 = -1, k1 = -1, k2 = -1, k3 = -1;           /* candidate SATDs (-1: not in the memo) */
            int m0 = 0, m1 = 0, m2 = 0, m3 = 0, c0 = -1, c1 = -1, c2 = -1, c3 = -1;   /* search result per candidate */
            {
                const int4* __restrict__ mp = memoRow + cx * MEMO_N;
#pragma unroll 1
                for (int k = 0; k < MEMO_N; k++)
                {
                    const int4 e = mp[k];
                    if (e.w < 0) break;                       /* entries are filled front to back */
                    if (e.x == nb0) { k0 = e.y; m0 = e.z; c0 = e.w; }
                    if (e.x == nb1) { k1 = e.y; m1 = e.z; c1 = e.w; }
                    if (e.x == nb2) { k2 = e.y; m2 = e.z; c2 = e.w; }
                    if (e.x == nb3) { k3 = e.y; m3 = e.z; c3 = e.w; }
                }
            }
            /* no candidates: the MVP is zero and no SATD is needed; otherwise every candidate needs its SATD */
            const bool miss = numc == 0 ? c0 < 0 : (k0 < 0 || (numc > 1 && k1 < 0) || (numc > 2 && k2 < 0) || (numc > 3 && k3 < 0));
            int outMv = 0, outCost = 0;
            if (!miss)
            {
                LaSearch s;
                la_search_begin(s, cx, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);
                la_upd_cand(s, k0, k1, k2, k3);
                const int mvp = la_pack_mv(s.mvpx, s.mvpy);
                /* the MVP is one of the candidates (or zero when there is none, which nb0 = 0 stands for) */
                const int rm = mvp == nb0 ? m0 : (mvp == nb1 ? m1 : (mvp == nb2 ? m2 : m3));
                const int rc = mvp == nb0 ? c0 : (mvp == nb1 ? c1 : (mvp == nb2 ? c2 : c3));
                s.outx = la_mv_x(rm); s.outy = la_mv_y(rm); s.outcost = rc;
                la_finish_skip(s);
                outMv = la_pack_mv(s.outx, s.outy); outCost = s.outcost;
            }
            /* ---- which lanes computed from true inputs only: the prefix up to the first wrong assumption / wait / miss ---- */
            const int rightMv = __shfl_up_sync(FULL_MASK, outMv, 1);
            const bool good = valid && avail && !miss && (lane == 0 || rightMv == r);
            const unsigned gm = __ballot_sync(FULL_MASK, good);
            const int n = gm == FULL_MASK ? 32 : __ffs(~gm) - 1;
            if (n > 0)
            {
                if (lane < n)
                {
                    const unsigned long long word = HAND_TAG | (uint32_t)outMv;
                    const int cuXY = cu + cuY * W;
                    myHand[cu] = word;
                    if (publishGlobal) myHandG[cu] = word;
                    mvMirror[cuXY] = outMv;
                    mcMirror[cuXY] = outCost;
                    mvOut[cuXY] = outMv;
                    mcOut[cuXY] = outCost;
                }
                prevMv = __shfl_sync(FULL_MASK, outMv, n - 1);
                x0 -= n;
                SSTAT_ADD(0, n);
                SSTAT_ADD(15, n);
                /* a short prefix means the estimate is poor here (or the row below is just ahead): go one CU at a time for a while */
                if (n < 4) serialLeft = 8;
                STRACE(cuY, x0 + n, 1, n);
                continue;
            }
        }
        else serialLeft--;

        /* ================= serial step: the whole warp works on CU x0 ================= */
        {
            const int cuX = x0;
            SSTAT_T(tp0);
            int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0, numc = 0;
            if (cuX < W - 1) { nb0 = prevMv; numc = 1; }
            if (!lastRow)
            {
                /* the row below runs right to left: its column cuX - 1 is published last */
                int bl = 0, br = 0;
                if (cuX > 0) bl = hand_wait(below + cuX - 1);
                const int mb = hand_wait(below + cuX);
                if (cuX < W - 1) br = hand_wait(below + cuX + 1);
                if (numc == 0) nb0 = mb; else nb1 = mb;
                numc++;
                if (cuX > 0) { if (numc == 1) nb1 = bl; else nb2 = bl; numc++; }
                if (cuX < W - 1) { if (numc == 2) nb2 = br; else nb3 = br; numc++; }
            }
            SSTAT_T(tp1);
            const int4 e = memoRow[cuX * MEMO_N + (lane & (MEMO_N - 1))];     /* lane k (mod MEMO_N) holds entry k */
            const int i0 = memo_find(e, nb0), i1 = memo_find(e, nb1), i2 = memo_find(e, nb2), i3 = memo_find(e, nb3);
            int k0 = __shfl_sync(FULL_MASK, e.y, i0 & 31), k1 = __shfl_sync(FULL_MASK, e.y, i1 & 31);
            int k2 = __shfl_sync(FULL_MASK, e.y, i2 & 31), k3 = __shfl_sync(FULL_MASK, e.y, i3 & 31);
            /* when every neighbour carries the same vector v the first candidate wins the strict-< chain whatever the costs;
             * its SATD only matters for the bidir skip rule, and only when v is zero (slicetype.cpp:2146-2149) */
            const bool allEq = (numc < 2 || nb1 == nb0) && (numc < 3 || nb2 == nb0) && (numc < 4 || nb3 == nb0);
            const bool ok = (allEq && !(nb0 == 0 && bidir && numc > 0)) ||
                            ((numc < 1 || (i0 >= 0 && k0 >= 0)) && (numc < 2 || (i1 >= 0 && k1 >= 0)) && (numc < 3 || (i2 >= 0 && k2 >= 0)) && (numc < 4 || (i3 >= 0 && k3 >= 0)));
            typename Px<P>::Row4 fe[4];
            SSTAT_T(tp2);
            if (!ok)
            {
                /* the reference's CAND pass, straight from global memory: quad k measures neighbour k */
                typename Px<P>::Row4 rr[4];
#pragma unroll
                for (int y = 0; y < 4; y++)
                    fe[y] = Px<P>::load_aligned(fencPlane + rowBase + 8 * cuX + y * stride);
                const int cmv = q == 0 ? nb0 : (q == 1 ? nb1 : (q == 2 ? nb2 : (q == 3 ? nb3 : 0)));
                fetch_qpel<P>(refPlane + rowBase + 8 * cuX, planeSize, stride, la_mv_x(cmv), la_mv_y(cmv), rr);
                const int cc = quad_sum(satd4x4_abs<P>(fe, rr)) >> 1;
                k0 = __shfl_sync(FULL_MASK, cc, 0); k1 = __shfl_sync(FULL_MASK, cc, 4);
                k2 = __shfl_sync(FULL_MASK, cc, 8); k3 = __shfl_sync(FULL_MASK, cc, 12);
                SSTAT_ADD(14, 1);
            }
            /* ---- the MVP, exactly as slicetype.cpp:2130-2150 ---- */
            SSTAT_T(tp3);
            LaSearch s;
            la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);
            la_upd_cand(s, k0, k1, k2, k3);
            const int mvp = la_pack_mv(s.mvpx, s.mvpy);
            const int im = memo_find(e, mvp);
            SSTAT_T(tp4);
            int rMv, rCost;
            if (im >= 0) { rMv = __shfl_sync(FULL_MASK, e.z, im); rCost = __shfl_sync(FULL_MASK, e.w, im); }
            else
            {
                /* nobody predicted this MVP: search it here, on the chain */
                if (ok)
                {
#pragma unroll
                    for (int y = 0; y < 4; y++)
                        fe[y] = Px<P>::load_aligned(fencPlane + rowBase + 8 * cuX + y * stride);
                }
                const LaneGeom L = lane_geom(lane, stride);
                int cs;
                bool cok;
                search_mv<P>(refPlane + 8 * cuY * stride + 8 * cuX, refPlane + rowBase + 8 * cuX, win, planeSize, stride, fe, lut, L, lane, bx, by,
                             cuX, cuY, W, H, mvp, rMv, rCost, cs, cok);
                SSTAT_ADD(1, 1);
            }
            SSTAT_T(tp5);
            SSTAT_ADD(20, tp1 - tp0); SSTAT_ADD(21, tp2 - tp1); SSTAT_ADD(22, tp3 - tp2); SSTAT_ADD(23, tp4 - tp3); SSTAT_ADD(24, tp5 - tp4);
            if (!ok) { SSTAT_ADD(25, tp3 - tp2); SSTAT_ADD(26, 1); }
            s.outx = la_mv_x(rMv); s.outy = la_mv_y(rMv); s.outcost = rCost;
            la_finish_skip(s);
            const int mvPacked = la_pack_mv(s.outx, s.outy);
            prevMv = mvPacked;
            if (lane == 0)
            {
                const unsigned long long word = HAND_TAG | (uint32_t)mvPacked;
                const int cuXY = cuX + cuY * W;
                myHand[cuX] = word;
                if (publishGlobal) myHandG[cuX] = word;
                mvMirror[cuXY] = mvPacked;
                mcMirror[cuXY] = s.outcost;
                mvOut[cuXY] = mvPacked;
                mcOut[cuXY] = s.outcost;
            }
            if (cuX > 0)
            {
                /* pull what the next CU of the chain is most likely to need into L1 while the result is being published:
                 * its source block and the 13-row windows of the four planes around this CU's MV */
                const int prow = lane & 15, pplane = lane >> 4;
                const P* w = refPlane + (8 * cuY + (la_mv_y(mvPacked) >> 2) - 2 + prow) * stride + 8 * (cuX - 1) + (la_mv_x(mvPacked) >> 2) - 4;
                if (prow < WIN_H)
                {
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(w + pplane * planeSize));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(w + pplane * planeSize + 20));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(w + (pplane + 2) * planeSize));
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(w + (pplane + 2) * planeSize + 20));
                }
                if (lane < 8)
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(fencPlane + (8 * cuY + lane) * stride + 8 * (cuX - 1)));
            }
            x0--;
            SSTAT_ADD(0, 1);
            SSTAT_ADD(17, 1);
            STRACE(cuY, x0 + 1, ok ? (im >= 0 ? 2 : 4) : (im >= 0 ? 3 : 5), 1);
        }
    }
}

#endif /* X265CU_SEARCH_CUH */
