/* x265cu_search.cuh -- the wavefront motion-search kernel (sm_100a).
 *
 * Replaces, for ONE reference list of one frame-cost estimate, the search half of
 * CostEstimateGroup::estimateCUCost (encoder/slicetype.cpp:2106-2160) and the lowres branch of
 * MotionEstimate::motionEstimate (encoder/motion.cpp:571-1172).  The L0 and L1 searches of an
 * estimate are independent of each other (they only meet in the bidir/intra decision, which
 * cost_kernel does afterwards for the whole frame in parallel), so the unit of work is
 * (job, list, cooperative slice), further cut into ROW GROUPS of a few CU rows: one small CTA per
 * group, one warp per CU row, so that the rows of one slice spread over many SMs and every warp
 * gets (close to) a scheduler of its own -- the search is a long dependent chain per CU, so
 * per-warp issue rate, not occupancy, is what sets the latency of an estimate.
 *
 * Rows form a wavefront because a CU's MVP candidates are its right, below, below-left and
 * below-right neighbours (slicetype.cpp:2117-2128): row y may process column x once row y+1
 * finished column max(x-1,0).  Inside a group the hand-off (progress counter + the row's MVs) goes
 * through shared memory; between groups it goes through global memory (L2): the top row of a
 * group publishes "CUs done" after a __threadfence, the bottom row of the group above polls it and
 * reads the MVs from the HBM mirror.  A group only ever waits for a group with a LOWER block
 * index (launched earlier), so the scheme cannot deadlock even when a launch does not fit the GPU.
 *
 * Inside a CU the search is a chain of dependent passes (la_core.h).  Each pass measures up to 8
 * candidate blocks at once: quad q = lane >> 2 owns candidate q, each lane its 4x4 sub-block; the
 * winner is one warp min-reduction over packed (cost << 3 | q) keys.  Control flow is warp-uniform.
 */
#ifndef X265CU_SEARCH_CUH
#define X265CU_SEARCH_CUH

struct SearchItem
{
    int job, list;
    int sliceFirstY, sliceLastY;   /* cooperative slice (or whole frame) this group belongs to */
    int firstY, lastY;             /* CU rows of this group (lastY = bottom row, processed first) */
    int progBase;                  /* index of row 0 of this (job, list) in the global progress array */
    int pad;
};

#define SEARCH_MAX_GROUP_ROWS 32

/* reference block sub-rows for this lane at quarter-pel MV (qx, qy); refLane = plane 0 at the lane's
 * 4x4 position of the current CU (lowres.h:62-103) */
template <typename P>
__device__ __forceinline__ void fetch_qpel(const P* __restrict__ refLane, int planeSize, int stride, int qx, int qy, typename Px<P>::Row4 out[4])
{
    const int hpelA = (qy & 2) | ((qx & 2) >> 1);
    const P* a = refLane + hpelA * planeSize + (qy >> 2) * stride + (qx >> 2);
    if ((qx | qy) & 1)
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const P* b = refLane + hpelB * planeSize + (qy2 >> 2) * stride + (qx2 >> 2);
#pragma unroll
        for (int i = 0; i < 4; i++)
            out[i] = Px<P>::avg(Px<P>::load(a + i * stride), Px<P>::load(b + i * stride));
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 4; i++)
            out[i] = Px<P>::load(a + i * stride);
    }
}

/* full-pel position (fx, fy): plane 0 only, no averaging */
template <typename P>
__device__ __forceinline__ void fetch_fpel(const P* __restrict__ refLane, int stride, int fx, int fy, typename Px<P>::Row4 out[4])
{
    const P* a = refLane + fy * stride + fx;
#pragma unroll
    for (int i = 0; i < 4; i++)
        out[i] = Px<P>::load(a + i * stride);
}

__device__ __forceinline__ uint32_t warp_min_key(bool valid, int cost, int q)
{
    uint32_t key = valid ? la_key(cost, q) : LA_KEY_NONE;
    return __reduce_min_sync(FULL_MASK, key);
}

template <typename P>
__global__ void __launch_bounds__(SEARCH_MAX_GROUP_ROWS * 32, 1)
search_kernel(const JobDev* __restrict__ jobs, const SearchItem* __restrict__ items, GeomDev g,
              const uint16_t* __restrict__ lut, int* gProg)
{
    extern __shared__ int sRing[];                 /* one row of W packed MVs per warp */
    __shared__ int sProg[SEARCH_MAX_GROUP_ROWS];   /* CUs done per row of this group */
    const SearchItem it = items[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nRows = it.lastY - it.firstY + 1;
    const int W = g.wCU, H = g.hCU;
    if (threadIdx.x < SEARCH_MAX_GROUP_ROWS) sProg[threadIdx.x] = 0;
    __syncthreads();
    if (warp >= nRows) return;

    const JobDev* __restrict__ jp = jobs + it.job;
    const int list = it.list;
    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int stride = g.stride, planeSize = (int)g.planeSize;
    const int bidir = jp->bidir;
    const P* __restrict__ fencPlane = (const P*)jp->fenc;
    const P* __restrict__ refPlane = (const P*)(list ? jp->ref1 : jp->ref0w);
    int* mvMirror = jp->mvs[list];
    int* __restrict__ mcMirror = jp->mvCosts[list];
    int* __restrict__ mvOut = jp->outMvs[list];
    int* __restrict__ mcOut = jp->outMvCosts[list];

    /* warp r owns row cuY = lastY - r; it depends on row cuY + 1 */
    const int cuY = it.lastY - warp;
    const bool lastRow = cuY == it.sliceLastY;              /* bottom row of the slice: no candidates from below */
    const bool belowIsGlobal = warp == 0 && !lastRow;       /* the row below belongs to another CTA */
    const bool publishGlobal = warp == nRows - 1 && cuY > it.sliceFirstY;
    volatile int* myRing = sRing + warp * W;
    volatile const int* belowRing = sRing + (warp > 0 ? warp - 1 : 0) * W;
    volatile const int* belowProg = belowIsGlobal ? (volatile const int*)(gProg + it.progBase + cuY + 1) : (volatile const int*)&sProg[warp > 0 ? warp - 1 : 0];
    volatile const int* belowMv = belowIsGlobal ? (volatile const int*)(mvMirror + (cuY + 1) * W) : belowRing;
    volatile int* myProgG = gProg + it.progBase + cuY;
    volatile int* myProgS = &sProg[warp];

    const int rowBase = (8 * cuY + by) * stride + bx;
    int prevMv = 0;                                /* MV of (cuX + 1, cuY): our own previous result */
    typename Px<P>::Row4 fe[4], feNext[4];
#pragma unroll
    for (int y = 0; y < 4; y++)
        feNext[y] = Px<P>::load_aligned(fencPlane + rowBase + 8 * (W - 1) + y * stride);

    for (int cuX = W - 1; cuX >= 0; cuX--)
    {
#pragma unroll
        for (int y = 0; y < 4; y++) fe[y] = feNext[y];
        if (cuX > 0)
        {
#pragma unroll
            for (int y = 0; y < 4; y++)
                feNext[y] = Px<P>::load_aligned(fencPlane + rowBase + 8 * (cuX - 1) + y * stride);
        }
        const P* __restrict__ refLane = refPlane + rowBase + 8 * cuX;

        /* ---- neighbour MVs (slicetype.cpp:2117-2128) ---- */
        int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0, numc = 0;
        if (cuX < W - 1) { nb0 = prevMv; numc = 1; }
        if (!lastRow)
        {
            const int needDone = W - (cuX > 0 ? cuX - 1 : 0);   /* row below finished column max(cuX-1, 0) */
            while (*belowProg < needDone) { }
            if (belowIsGlobal) __threadfence(); else __threadfence_block();
            const int mb = belowMv[cuX];
            if (numc == 0) nb0 = mb; else nb1 = mb;
            numc++;
            if (cuX > 0)
            {
                const int bl = belowMv[cuX - 1];
                if (numc == 1) nb1 = bl; else nb2 = bl;
                numc++;
            }
            if (cuX < W - 1)
            {
                const int br = belowMv[cuX + 1];
                if (numc == 2) nb2 = br; else nb3 = br;
                numc++;
            }
        }
        LaSearch s;
        la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);

        /* ---- CAND: SATD at each neighbour MV, no mvcost ---- */
        if (numc)
        {
            const int p = la_cand_mv(s, q & 3);
            int part = 0;
            if (q < numc)
            {
                typename Px<P>::Row4 r[4];
                fetch_qpel<P>(refLane, planeSize, stride, la_mv_x(p), la_mv_y(p), r);
                part = satd4x4_abs<P>(fe, r);
            }
            const int cost = quad_sum(part) >> 1;
            la_upd_cand(s, __shfl_sync(FULL_MASK, cost, 0), __shfl_sync(FULL_MASK, cost, 4),
                        __shfl_sync(FULL_MASK, cost, 8), __shfl_sync(FULL_MASK, cost, 12));
        }
        const uint16_t* __restrict__ lutx = lut - s.mvpx;
        const uint16_t* __restrict__ luty = lut - s.mvpy;

        /* ---- START: qpel MVP (no mvcost) / rounded MVP / zero ---- */
        la_enter_start(s);
        {
            const bool sp = la_start_subpel(s), nz = la_start_nonzero(s);
            /* one code path for the three candidates: q0 = (pmx, pmy), q1 = rounded, q2 = zero */
            const int qx = q == 0 ? s.pmx : (q == 1 ? ((s.pmx + 2) >> 2) * 4 : 0);
            const int qy = q == 0 ? s.pmy : (q == 1 ? ((s.pmy + 2) >> 2) * 4 : 0);
            const bool valid = q == 0 || (q == 1 && sp) || (q == 2 && nz);
            int part = 0;
            if (valid)
            {
                typename Px<P>::Row4 r[4];
                fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                part = sad4x4<P>(fe, r);
            }
            const int mvc = q == 0 ? 0 : lutx[qx] + luty[qy];
            const int cost = quad_sum(part) + mvc;
            la_upd_start(s, __shfl_sync(FULL_MASK, cost, 0), __shfl_sync(FULL_MASK, cost, 4), __shfl_sync(FULL_MASK, cost, 8));
        }

        /* ---- HEX6 + HEX3 rounds: full-pel SAD + mvcost ---- */
        {
            const int fx = s.bmx + la_hex2x((q + 1) & 7), fy = s.bmy + la_hex2y((q + 1) & 7);
            int part = 0;
            if (q < 6)
            {
                typename Px<P>::Row4 r[4];
                fetch_fpel<P>(refLane, stride, fx, fy, r);
                part = sad4x4<P>(fe, r);
            }
            const int cost = quad_sum(part) + lutx[fx * 4] + luty[fy * 4];
            bool more = la_upd_hex6(s, warp_min_key(q < 6, cost, q));
            while (more)
            {
                const int hx = s.bmx + la_hex2x((s.dir + q) & 7), hy = s.bmy + la_hex2y((s.dir + q) & 7);
                int p3 = 0;
                if (q < 3)
                {
                    typename Px<P>::Row4 r[4];
                    fetch_fpel<P>(refLane, stride, hx, hy, r);
                    p3 = sad4x4<P>(fe, r);
                }
                const int c3 = quad_sum(p3) + lutx[hx * 4] + luty[hy * 4];
                more = la_upd_hex3(s, warp_min_key(q < 3, c3, q));
            }
        }

        /* ---- SQ8: 8-point square ---- */
        bool subpel;
        {
            const int fx = s.bmx + la_sq1x(q + 1), fy = s.bmy + la_sq1y(q + 1);
            typename Px<P>::Row4 r[4];
            fetch_fpel<P>(refLane, stride, fx, fy, r);
            const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
            subpel = la_upd_sq8(s, warp_min_key(true, cost, q), lut);
        }

        if (subpel)
        {
            /* ---- HPEL: 4 half-pel SADs ---- */
            {
                const int qx = s.bmx + la_sq1x((q + 1) & 7) * 2, qy = s.bmy + la_sq1y((q + 1) & 7) * 2;
                int part = 0;
                if (q < 4)
                {
                    typename Px<P>::Row4 r[4];
                    fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                    part = sad4x4<P>(fe, r);
                }
                const int cost = quad_sum(part) + lutx[qx] + luty[qy];
                la_upd_hpel(s, warp_min_key(q < 4, cost, q));
            }
            /* ---- QPEL: SATD re-measure (q0) + 4 quarter-pel SATDs ---- */
            {
                const int qx = s.bmx + la_sq1x(q), qy = s.bmy + la_sq1y(q);
                int part = 0;
                if (q < 5)
                {
                    typename Px<P>::Row4 r[4];
                    fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                    part = satd4x4_abs<P>(fe, r);
                }
                const int cost = (quad_sum(part) >> 1) + lutx[qx] + luty[qy];
                const int c0 = __shfl_sync(FULL_MASK, cost, 0);
                la_upd_qpel(s, c0, warp_min_key(q >= 1 && q < 5, cost, q));
            }
        }
        la_finish_skip(s);

        const int mvPacked = la_pack_mv(s.outx, s.outy);
        prevMv = mvPacked;
        if (lane == 0)
        {
            const int cuXY = cuX + cuY * W;
            myRing[cuX] = mvPacked;
            ((volatile int*)mvMirror)[cuXY] = mvPacked;
            mcMirror[cuXY] = s.outcost;
            mvOut[cuXY] = mvPacked;
            mcOut[cuXY] = s.outcost;
            __threadfence_block();
            *myProgS = W - cuX;
            if (publishGlobal)
            {
                __threadfence();
                *myProgG = W - cuX;
            }
        }
    }
}

#endif /* X265CU_SEARCH_CUH */
