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
 * below-right neighbours (slicetype.cpp:2117-2128).  A finished CU publishes ONE 64-bit word
 * {tag = 1, packed MV}; the row above polls exactly the words it needs.  Because tag and data
 * travel in the same naturally aligned 8-byte store there is no separate progress counter and no
 * memory fence anywhere on the chain.  Inside a group the words live in shared memory, between
 * groups in a global hand-off row (L2).  A group only ever waits for a group with a LOWER block
 * index (launched earlier), so the scheme cannot deadlock even when a launch does not fit the GPU.
 *
 * Inside a CU the search is a chain of dependent passes (la_core.h).  Each pass measures up to 8
 * candidate blocks at once: quad q = lane >> 2 owns candidate q, each lane its 4x4 sub-block; the
 * winner is one warp min-reduction over packed (cost << 3 | q) keys.  Passes are branch-free:
 * every quad always measures a (valid-address) block and invalid candidates are masked out of the
 * key reduction; the per-lane candidate offsets of every pass are computed once per kernel.
 */
#ifndef X265CU_SEARCH_CUH
#define X265CU_SEARCH_CUH

struct SearchItem
{
    int job, list;
    int sliceFirstY, sliceLastY;   /* cooperative slice (or whole frame) this group belongs to */
    int firstY, lastY;             /* CU rows of this group (lastY = bottom row, processed first) */
    int pubBase;                   /* hand-off row this group's TOP row publishes to (index of entry 0), -1: none */
    int subBase;                   /* hand-off row the group's BOTTOM row reads (the group below), -1: none */
};

#define SEARCH_MAX_GROUP_ROWS 8
#define HAND_TAG (1ull << 32)

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

/* wait for a published hand-off word and return its MV */
__device__ __forceinline__ int hand_wait(volatile const unsigned long long* e)
{
    unsigned long long v;
    do { v = *e; } while (!(v & HAND_TAG));
    return (int)(uint32_t)v;
}

template <typename P>
__global__ void __launch_bounds__(SEARCH_MAX_GROUP_ROWS * 32, 3)
search_kernel(const JobDev* __restrict__ jobs, const SearchItem* __restrict__ items, GeomDev g,
              const uint16_t* __restrict__ lut, unsigned long long* gHand)
{
    extern __shared__ unsigned long long sHand[];  /* [rows of the group][W] hand-off words */
    const SearchItem it = items[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nRows = it.lastY - it.firstY + 1;
    const int W = g.wCU, H = g.hCU;
    for (int i = threadIdx.x; i < nRows * W; i += blockDim.x) sHand[i] = 0;
    const JobDev* __restrict__ jp = jobs + it.job;
    const int list = it.list;
    {
        /* Stream this group's band of the source plane and of the four reference planes into L2
         * before the dependent chain starts: a single estimate reads frames that left L2 long
         * ago, and a DRAM miss inside a pass stalls the whole chain for ~1 us.  Fire and forget. */
        const int bandRows = (it.lastY - it.firstY + 1) * 8 + 64;           /* +-32 rows of search range */
        const int bandTop = it.firstY * 8 - 32;
        const int linesPerRow = (g.width + 64) * (int)sizeof(P) / 128 + 1;
        const char* fencB = (const char*)jp->fenc;
        const char* refB = (const char*)(list ? jp->ref1 : jp->ref0w);
        const int total = bandRows * linesPerRow * 5;
        for (int i = threadIdx.x; i < total; i += blockDim.x)
        {
            const int plane = i / (bandRows * linesPerRow);                  /* 0..3 reference planes, 4 = source */
            const int rem = i - plane * bandRows * linesPerRow;
            const int row = bandTop + rem / linesPerRow, line = rem % linesPerRow;
            if (plane == 4 && (row < it.firstY * 8 || row >= (it.lastY + 1) * 8)) continue;
            const char* base = plane == 4 ? fencB : refB + (int64_t)plane * g.planeSize * (int)sizeof(P);
            const char* ptr = base + ((int64_t)row * g.stride - 32) * (int)sizeof(P) + line * 128;
            asm volatile("prefetch.global.L2 [%0];" :: "l"(ptr));
        }
    }
    __syncthreads();
    if (warp >= nRows) return;

    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int stride = g.stride, planeSize = (int)g.planeSize;
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

    /* per-lane candidate geometry of the fixed-shape passes (motion.cpp:64-66 tables) */
    const int hex6dx = la_hex2x((q + 1) & 7), hex6dy = la_hex2y((q + 1) & 7);
    const int hex6off = hex6dy * stride + hex6dx;
    const int sq8dx = la_sq1x(q + 1), sq8dy = la_sq1y(q + 1);
    const int sq8off = sq8dy * stride + sq8dx;
    const int hpdx = la_sq1x((q + 1) & 7) * 2, hpdy = la_sq1y((q + 1) & 7) * 2;   /* quarter-pel units */
    const int qpdx = la_sq1x(q), qpdy = la_sq1y(q);

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
        if (cuX > 0)
        {
            /* pull the likely window of the NEXT CU (same MV as our right neighbour, 8 samples to the
             * left) into L1 while this CU is being searched: 16 rows x 4 planes, two sectors per row */
            const int prow = (lane & 15) - 4, pplane = lane >> 4;
            const P* w = refPlane + (8 * cuY + (la_mv_y(prevMv) >> 2) + prow) * stride + 8 * (cuX - 1) + (la_mv_x(prevMv) >> 2) - 4;
            asm volatile("prefetch.global.L1 [%0];" :: "l"(w + pplane * planeSize));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(w + pplane * planeSize + 16));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(w + (pplane + 2) * planeSize));
            asm volatile("prefetch.global.L1 [%0];" :: "l"(w + (pplane + 2) * planeSize + 16));
        }

        /* ---- neighbour MVs (slicetype.cpp:2117-2128): right, below, below-left, below-right ---- */
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
        LaSearch s;
        la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);

        /* ---- CAND: SATD at each neighbour MV, no mvcost (quads >= numc re-measure candidate 0) ---- */
        if (numc)
        {
            const int p = la_cand_mv(s, q < numc ? q : 0);
            typename Px<P>::Row4 r[4];
            fetch_qpel<P>(refLane, planeSize, stride, la_mv_x(p), la_mv_y(p), r);
            const int cost = quad_sum(satd4x4_abs<P>(fe, r)) >> 1;
            la_upd_cand(s, __shfl_sync(FULL_MASK, cost, 0), __shfl_sync(FULL_MASK, cost, 4),
                        __shfl_sync(FULL_MASK, cost, 8), __shfl_sync(FULL_MASK, cost, 12));
        }
        const uint16_t* __restrict__ lutx = lut - s.mvpx;
        const uint16_t* __restrict__ luty = lut - s.mvpy;

        /* ---- START: q0 = qpel MVP (no mvcost), q1 = rounded MVP, q2 = zero ---- */
        la_enter_start(s);
        {
            const int qx = q == 0 ? s.pmx : (q == 1 ? ((s.pmx + 2) >> 2) * 4 : 0);
            const int qy = q == 0 ? s.pmy : (q == 1 ? ((s.pmy + 2) >> 2) * 4 : 0);
            typename Px<P>::Row4 r[4];
            fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
            const int mvc = q == 0 ? 0 : lutx[qx] + luty[qy];
            const int cost = quad_sum(sad4x4<P>(fe, r)) + mvc;
            la_upd_start(s, __shfl_sync(FULL_MASK, cost, 0), __shfl_sync(FULL_MASK, cost, 4), __shfl_sync(FULL_MASK, cost, 8));
        }

        /* ---- HEX6 + HEX3 rounds: full-pel SAD + mvcost ---- */
        {
            const int fx = s.bmx + hex6dx, fy = s.bmy + hex6dy;
            typename Px<P>::Row4 r[4];
            fetch_off<P>(refLane, stride, s.bmy * stride + s.bmx + hex6off, r);
            const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
            bool more = la_upd_hex6(s, warp_min_key(q < 6, cost, q));
            while (more)
            {
                const int hdx = la_hex2x((s.dir + q) & 7), hdy = la_hex2y((s.dir + q) & 7);
                const int hx = s.bmx + hdx, hy = s.bmy + hdy;
                typename Px<P>::Row4 r3[4];
                fetch_off<P>(refLane, stride, hy * stride + hx, r3);
                const int c3 = quad_sum(sad4x4<P>(fe, r3)) + lutx[hx * 4] + luty[hy * 4];
                more = la_upd_hex3(s, warp_min_key(q < 3, c3, q));
            }
        }

        /* ---- SQ8: 8-point square ---- */
        bool subpel;
        {
            const int fx = s.bmx + sq8dx, fy = s.bmy + sq8dy;
            typename Px<P>::Row4 r[4];
            fetch_off<P>(refLane, stride, s.bmy * stride + s.bmx + sq8off, r);
            const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
            subpel = la_upd_sq8(s, warp_min_key(true, cost, q), lut);
        }

        if (subpel)
        {
            /* ---- HPEL: 4 half-pel SADs ---- */
            {
                const int qx = s.bmx + hpdx, qy = s.bmy + hpdy;
                typename Px<P>::Row4 r[4];
                fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[qx] + luty[qy];
                la_upd_hpel(s, warp_min_key(q < 4, cost, q));
            }
            /* ---- QPEL: SATD re-measure (q0) + 4 quarter-pel SATDs ---- */
            {
                const int qx = s.bmx + qpdx, qy = s.bmy + qpdy;
                typename Px<P>::Row4 r[4];
                fetch_qpel<P>(refLane, planeSize, stride, qx, qy, r);
                const int cost = (quad_sum(satd4x4_abs<P>(fe, r)) >> 1) + lutx[qx] + luty[qy];
                const int c0 = __shfl_sync(FULL_MASK, cost, 0);
                la_upd_qpel(s, c0, warp_min_key(q >= 1 && q < 5, cost, q));
            }
        }
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
    }
}

#endif /* X265CU_SEARCH_CUH */
