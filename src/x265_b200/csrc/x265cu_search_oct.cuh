/* x265cu_search_oct.cuh -- the wavefront motion-search kernel, octet form (sm_100a).
 *
 * Same search as x265cu_search_plain.cuh (one reference list of one estimate: encoder/slicetype.cpp:2106-2160 + the
 * lowres branch of MotionEstimate::motionEstimate, encoder/motion.cpp:571-1172; decisions by la_core.h), other mapping.
 * ncu showed the quad form issue-bound with ~1200 warp-instructions per CU of which a few per cent are SAD/SATD math: a
 * warp measures <= 8 candidates of ONE CU per pass, each lane a 4x4, so every lane's useful work is ~12 instructions
 * between ~150 of addressing, reductions and the (redundant, warp-uniform) decision logic.  Here
 *
 *   - an OCTET (8 lanes) owns a CU; a warp runs FOUR CU rows at once, skewed by two columns per row (the wavefront
 *     dependency: right, below, below-left, below-right), in lock step: the neighbour vectors of the row below are the
 *     last three results of the octet next door and travel by shuffle, nothing is waited for inside a warp;
 *   - SAD passes: lane r owns pixel ROW r of the 8x8 block and loops over the pass's candidates (8 samples per
 *     __vsadu4 pair), the per-row sums of all candidates are reduced over the octet by one transposing butterfly
 *     (7 shuffles), lane k ends up with candidate k's cost, adds ITS mvcost and the winner is one redux.min over the
 *     octet's packed (cost << 3 | k) keys -- the reference's COPYn_IF_LT packing (motion.cpp:693-725);
 *   - SATD passes (neighbour candidates, quarter-pel refine): lane = 4x4 sub-block as before, two candidates per round;
 *   - the decision logic runs once per octet instead of once per warp: a quarter of the redundant instructions.
 *
 * Only every fourth row boundary needs a hand-off word (tagged 64-bit {1, MV}, shared memory between the warps of a
 * CTA, L2 between CTAs).  CTAs take their work item from an atomic ticket, so an item only ever waits for an item that
 * a running CTA has already acquired: no assumption about the order in which the hardware starts blocks.
 */
#ifndef X265CU_SEARCH_OCT_CUH
#define X265CU_SEARCH_OCT_CUH

#define OCT_MAX_WARPS 8          /* bands of 4 CU rows per CTA */
#ifndef OCT_MIN_CTAS
#define OCT_MIN_CTAS 4           /* x 256 threads: 64 registers, 32 warps per SM (the kernel is latency-bound per warp) */
#endif
#define OCT_FENC_UNITS 16        /* the CU's source block in shared memory: 8 rows x 2 units of 4 samples */
#define HAND_SPIN_LIMIT (1u << 25)

/* wait for a published hand-off word; a wait that does not end (it cannot: see the ticket) raises the error word */
__device__ __forceinline__ int hand_wait_safe(volatile const unsigned long long* e, unsigned int* err, unsigned sleepNs)
{
    unsigned long long v = *e;
    unsigned int spins = 0;
    while (!(v & HAND_TAG))
    {
        if (++spins > HAND_SPIN_LIMIT) { *err = 1; return 0; }
        __nanosleep(sleepNs);       /* the band below needs microseconds per step: do not burn the issue slots it could use */
        v = *e;
    }
    return (int)(uint32_t)v;
}

template <typename P> struct Row8 { typename Px<P>::Row4 lo, hi; };

/* 8 samples of window row y starting x (0..8) samples into the row, plane `plane` */
template <typename P>
__device__ __forceinline__ Row8<P> win_row8(const typename Px<P>::Row4* win, int plane, int x, int y)
{
    const typename Px<P>::Row4* u = win + plane * WIN_PLANE_UNITS + y * WIN_ROW_UNITS + (x >> 2);
    const typename Px<P>::Row4 u0 = u[0], u1 = u[1], u2 = u[2];
    Row8<P> o;
    o.lo = Px<P>::combine(u0, u1, x & 3);
    o.hi = Px<P>::combine(u1, u2, x & 3);
    return o;
}

template <typename P>
__device__ __forceinline__ Row8<P> glob_row8(const P* __restrict__ p)
{
    Row8<P> o;
    o.lo = Px<P>::load(p);
    o.hi = Px<P>::load(p + 4);
    return o;
}

/* ---- row r of the 8x8 reference block displaced by a vector: lowresMC / lowresQPelCost (common/lowres.h:62-103) ----
 * Fast forms read the octet's window (the caller has checked that the whole pass lies inside it); the slow forms read
 * global memory, are rare (a search that walked out of its window) and are kept out of line: the kernel is
 * instruction-cache bound before it is anything else, every inlined copy of a rare path costs the common one. */
template <typename P>
__device__ __forceinline__ Row8<P> win_row_qpel(const typename Px<P>::Row4* win, int wx0, int wy0, int r, int qx, int qy)
{
    const int hpelA = (qy & 2) | ((qx & 2) >> 1);
    Row8<P> a = win_row8<P>(win, hpelA, (qx >> 2) - wx0, (qy >> 2) - wy0 + r);
    if ((qx | qy) & 1)                                  /* uniform over the octet */
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const Row8<P> b = win_row8<P>(win, hpelB, (qx2 >> 2) - wx0, (qy2 >> 2) - wy0 + r);
        a.lo = Px<P>::avg(a.lo, b.lo); a.hi = Px<P>::avg(a.hi, b.hi);
    }
    return a;
}

/* SAD of row r at quarter-pel vector (qx, qy) from global memory; refRow = plane 0 at (8 cuX, 8 cuY + r) */
template <typename P>
__device__ __noinline__ int glob_sad_qpel(const P* __restrict__ refRow, int planeSize, int stride, typename Px<P>::Row4 flo, typename Px<P>::Row4 fhi, int qx, int qy)
{
    const int hpelA = (qy & 2) | ((qx & 2) >> 1);
    Row8<P> a = glob_row8<P>(refRow + (hpelA * planeSize + (qy >> 2) * stride + (qx >> 2)));
    if ((qx | qy) & 1)
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const Row8<P> b = glob_row8<P>(refRow + (hpelB * planeSize + (qy2 >> 2) * stride + (qx2 >> 2)));
        a.lo = Px<P>::avg(a.lo, b.lo); a.hi = Px<P>::avg(a.hi, b.hi);
    }
    return Px<P>::sad(flo, a.lo) + Px<P>::sad(fhi, a.hi);
}

/* is the 8x8 block at quarter-pel vectors (qx +- 2, qy +- 2) (both sources of every one of them) inside the window? */
__device__ __forceinline__ bool win_holds_qpel(int wx0, int wy0, int qx, int qy, int reach)
{
    return (unsigned)(((qx - reach) >> 2) - wx0) <= (unsigned)(WIN_W - 8) && (unsigned)(((qx + reach + 1) >> 2) - wx0) <= (unsigned)(WIN_W - 8) &&
           (unsigned)(((qy - reach) >> 2) - wy0) <= (unsigned)(WIN_H - 8) && (unsigned)(((qy + reach + 1) >> 2) - wy0) <= (unsigned)(WIN_H - 8);
}
__device__ __forceinline__ bool win_holds_fpel(int wx0, int wy0, int fx, int fy, int reach)
{
    return (unsigned)(fx - reach - wx0) <= (unsigned)(WIN_W - 8) && (unsigned)(fx + reach - wx0) <= (unsigned)(WIN_W - 8) &&
           (unsigned)(fy - reach - wy0) <= (unsigned)(WIN_H - 8) && (unsigned)(fy + reach - wy0) <= (unsigned)(WIN_H - 8);
}

/* candidate k of a pass, as a quarter-pel vector (motion.cpp:64-66 tables; the passes of la_core.h) */
enum { PASS_HEX6 = 0, PASS_HEX3, PASS_SQ8, PASS_HPEL };
__device__ __forceinline__ void pass_cand(int kind, int k, int bmx, int bmy, int dir, int& qx, int& qy)
{
    if (kind == PASS_HEX6) { qx = (bmx + la_hex2x(k + 1)) * 4; qy = (bmy + la_hex2y(k + 1)) * 4; }
    else if (kind == PASS_HEX3) { qx = (bmx + la_hex2x((dir + k) & 7)) * 4; qy = (bmy + la_hex2y((dir + k) & 7)) * 4; }
    else if (kind == PASS_SQ8) { qx = (bmx + la_sq1x(k + 1)) * 4; qy = (bmy + la_sq1y(k + 1)) * 4; }
    else { qx = bmx + la_sq1x(k + 1) * 2; qy = bmy + la_sq1y(k + 1) * 2; }
}

/* a whole SAD pass from global memory (the pass left the window): lane r gets candidate r's sum.  Rare, out of line. */
template <typename P>
__device__ __noinline__ int oct_pass_slow(const P* __restrict__ refRow, int planeSize, int stride, typename Px<P>::Row4 flo, typename Px<P>::Row4 fhi,
                                          int kind, int n, int bmx, int bmy, int dir, int r, unsigned gmask)
{
    int tot = 0;
    for (int k = 0; k < n; k++)
    {
        int qx, qy;
        pass_cand(kind, k, bmx, bmy, dir, qx, qy);
        int v = glob_sad_qpel<P>(refRow, planeSize, stride, flo, fhi, qx, qy);
        v += __shfl_xor_sync(gmask, v, 1);
        v += __shfl_xor_sync(gmask, v, 2);
        v += __shfl_xor_sync(gmask, v, 4);
        if (k == r) tot = v;
    }
    return tot;
}

template <typename P>
__device__ __forceinline__ int sad_row8(const Row8<P>& f, const Row8<P>& r) { return Px<P>::sad(f.lo, r.lo) + Px<P>::sad(f.hi, r.hi); }

/* v[k] = this lane's (row's) part of candidate k; returns candidate (lane & 7)'s sum over the 8 lanes of the octet */
__device__ __forceinline__ int oct_reduce8(const int v[8], int r, unsigned gmask)
{
    const bool b2 = (r & 4) != 0, b1 = (r & 2) != 0, b0 = (r & 1) != 0;
    int w[4];
#pragma unroll
    for (int j = 0; j < 4; j++)
    {
        const int keep = b2 ? v[j + 4] : v[j], send = b2 ? v[j] : v[j + 4];
        w[j] = keep + __shfl_xor_sync(gmask, send, 4);
    }
    int x[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
        const int keep = b1 ? w[j + 2] : w[j], send = b1 ? w[j] : w[j + 2];
        x[j] = keep + __shfl_xor_sync(gmask, send, 2);
    }
    const int keep = b0 ? x[1] : x[0], send = b0 ? x[0] : x[1];
    return keep + __shfl_xor_sync(gmask, send, 1);
}

/* the same for <= 4 candidates: lane r ends up with candidate (r & 3) */
__device__ __forceinline__ int oct_reduce4(const int v[4], int r, unsigned gmask)
{
    const bool b1 = (r & 2) != 0, b0 = (r & 1) != 0;
    int w[4];
#pragma unroll
    for (int j = 0; j < 4; j++) w[j] = v[j] + __shfl_xor_sync(gmask, v[j], 4);
    int x[2];
#pragma unroll
    for (int j = 0; j < 2; j++)
    {
        const int keep = b1 ? w[j + 2] : w[j], send = b1 ? w[j] : w[j + 2];
        x[j] = keep + __shfl_xor_sync(gmask, send, 2);
    }
    const int keep = b0 ? x[1] : x[0], send = b0 ? x[0] : x[1];
    return keep + __shfl_xor_sync(gmask, send, 1);
}

__device__ __forceinline__ int quad_sum_m(int v, unsigned gmask)
{
    v += __shfl_xor_sync(gmask, v, 1);
    v += __shfl_xor_sync(gmask, v, 2);
    return v;
}

/* SATD of the 8x8 block at quarter-pel vector (qx, qy): this lane's 4x4 sub-block (bx, by) against fe[]; not yet summed */
template <typename P>
__device__ __noinline__ int glob_satd4x4(const P* __restrict__ refLane, int planeSize, int stride, typename Px<P>::Row4 f0, typename Px<P>::Row4 f1,
                                         typename Px<P>::Row4 f2, typename Px<P>::Row4 f3, int qx, int qy)
{
    typename Px<P>::Row4 fe[4] = { f0, f1, f2, f3 }, rr[4];
    fetch_qpel<P>(refLane, planeSize, stride, qx, qy, rr);
    return satd4x4_abs<P>(fe, rr);
}

template <typename P, bool WIN>
__device__ __forceinline__ int oct_satd(const P* __restrict__ refCU, int planeSize, int stride, const typename Px<P>::Row4* win,
                                        const typename Px<P>::Row4* fblk, int wx0, int wy0, int bx, int by, int qx, int qy, bool inWin, unsigned gmask)
{
    typename Px<P>::Row4 fe[4], rr[4];
#pragma unroll
    for (int i = 0; i < 4; i++) fe[i] = fblk[(by + i) * 2 + (bx >> 2)];
    int v;
    if (WIN && inWin)
    {
        win_qpel<P>(win, bx - wx0, by - wy0, qx, qy, rr);
        v = satd4x4_abs<P>(fe, rr);
    }
    else
        v = glob_satd4x4<P>(refCU + by * stride + bx, planeSize, stride, fe[0], fe[1], fe[2], fe[3], qx, qy);
    return quad_sum_m(v, gmask) >> 1;
}

template <typename P, bool WIN>
__global__ void __launch_bounds__(OCT_MAX_WARPS * 32, OCT_MIN_CTAS)
oct_search_kernel(const JobDev* __restrict__ jobs, const SearchPlan* __restrict__ plans, const SearchItem* __restrict__ items, GeomDev g,
                  const uint16_t* __restrict__ lut, unsigned long long* gHand, SearchCtl* ctl, int slack, unsigned sleepNs)
{
    typedef typename Px<P>::Row4 R4;
    extern __shared__ unsigned long long sHand[];   /* [bands][W] hand-off words, [bands * 4] windows, [bands * 4] source blocks */
    const SearchItem it = items[take_ticket(&ctl->ticket[0])];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int maxBands = blockDim.x >> 5;
    const int nRows = it.lastY - it.firstY + 1, nBands = (nRows + 3) >> 2;
    const int W = g.wCU, H = g.hCU;
    R4* winAll = (R4*)(sHand + maxBands * W);
    R4* fblkAll = winAll + maxBands * 4 * WIN_PITCH;
    for (int i = threadIdx.x; i < nBands * W; i += blockDim.x) sHand[i] = 0;
    const SearchPlan pl = plans[it.search];
    const JobDev* __restrict__ jp = jobs + pl.job;
    const int list = pl.list;
    {
        /* stream this item's band of the source plane and of the four reference planes into L2 before the dependent
         * chain starts (a DRAM miss inside a pass stalls the whole chain).  Fire and forget. */
        const int bandRows = nRows * 8 + 64;                                /* +-32 rows of search range */
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
    if (warp >= nBands) return;

    const int g4 = lane >> 3, r = lane & 7;
    const unsigned gmask = 0xFFu << (8 * g4);
    const int sub = r & 3, half = r >> 2, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int stride = g.stride, planeSize = (int)g.planeSize;
    const int bidir = jp->bidir;
    const P* __restrict__ fencPlane = (const P*)jp->fenc;
    const P* __restrict__ refPlane = (const P*)(list ? jp->ref1 : jp->ref0w);
    int* __restrict__ mvMirror = jp->mvs[list];
    int* __restrict__ mcMirror = jp->mvCosts[list];
    int* __restrict__ mvOut = jp->outMvs[list];
    int* __restrict__ mcOut = jp->outMvCosts[list];

    /* octet g4 of band `warp` owns row cuY; it depends on row cuY + 1 = the octet next door, or the band below */
    const int bandRows = nRows - 4 * warp < 4 ? nRows - 4 * warp : 4;
    const int cuY = it.lastY - 4 * warp - g4;
    const bool rowActive = g4 < bandRows;
    const bool lastRow = cuY == it.sliceLastY;              /* bottom row of the slice: no candidates from below */
    /* the band's top row is what the band above reads */
    const bool publisher = g4 == bandRows - 1;
    volatile unsigned long long* pubHand = NULL;
    if (publisher)
    {
        if (warp < nBands - 1) pubHand = sHand + warp * W;
        else if (it.pubBase >= 0) pubHand = gHand + it.pubBase;
    }
    volatile const unsigned long long* below = (warp == 0) ? (volatile const unsigned long long*)(gHand + (it.subBase >= 0 ? it.subBase : 0))
                                                           : (volatile const unsigned long long*)(sHand + (warp - 1) * W);
    const bool band0HasBelow = it.lastY - 4 * warp != it.sliceLastY;
    R4* win = winAll + (warp * 4 + g4) * WIN_PITCH;
    R4* fblk = fblkAll + (warp * 4 + g4) * OCT_FENC_UNITS;

    const P* __restrict__ fencRow = fencPlane + (8 * (rowActive ? cuY : it.lastY) + r) * stride;
    const P* __restrict__ refRow0 = refPlane + (8 * (rowActive ? cuY : it.lastY) + r) * stride;
    const P* __restrict__ refCU0 = refPlane + (8 * (rowActive ? cuY : it.lastY)) * stride;
    int h1 = 0, h2 = 0, h3 = 0;                    /* this row's last three results (columns x + 1, x + 2, x + 3) */
    const int srcBelow = ((g4 + 3) & 3) * 8;       /* a lane of the octet that owns the row below */

    for (int t = 0; t < W + 6; t++)
    {
        const int cuX = W - 1 - (t - 2 * g4);
        const bool active = rowActive && cuX >= 0 && cuX < W;
        /* ---- neighbour MVs of the row below (slicetype.cpp:2117-2128): column cuX - 1 is the newest result there ---- */
        int bl = __shfl_sync(FULL_MASK, h1, srcBelow);
        int mb = __shfl_sync(FULL_MASK, h2, srcBelow);
        int br = __shfl_sync(FULL_MASK, h3, srcBelow);
        const int x0 = W - 1 - t;                   /* column of octet 0 */
        if (band0HasBelow && x0 >= 0)
        {
            /* octet 0 reads the band below through hand-off words; the whole warp waits (it runs in lock step) */
            /* the row below runs right to left: once its newest word is there, the two older ones are */
            /* `slack` > 0 (launches that fill the GPU): stay that many columns further behind the band below, so that the
             * jitter of its steps is absorbed by the distance instead of by a wait in every step */
            int wbl = 0, wbr = 0;
            const int xw = x0 - 1 - slack > 0 ? x0 - 1 - slack : 0;
            hand_wait_safe(below + xw, &ctl->error, sleepNs);
            if (x0 > 0) wbl = (int)(uint32_t)below[x0 - 1];
            const int wmb = (int)(uint32_t)below[x0];
            if (x0 < W - 1) wbr = (int)(uint32_t)below[x0 + 1];
            if (g4 == 0) { bl = wbl; mb = wmb; br = wbr; }
        }
        int result = 0;
        if (active)
        {
            const int prevMv = h1;                  /* MV of (cuX + 1, cuY): our own previous result */
            const P* __restrict__ refRow = refRow0 + 8 * cuX;
            const P* __restrict__ refCU = refCU0 + 8 * cuX;
            Row8<P> fe;
            fe.lo = Px<P>::load_aligned(fencRow + 8 * cuX);
            fe.hi = Px<P>::load_aligned(fencRow + 8 * cuX + 4);

            int nb0 = 0, nb1 = 0, nb2 = 0, nb3 = 0, numc = 0;
            if (cuX < W - 1) { nb0 = prevMv; numc = 1; }
            if (!lastRow)
            {
                if (numc == 0) nb0 = mb; else nb1 = mb;
                numc++;
                if (cuX > 0) { if (numc == 1) nb1 = bl; else nb2 = bl; numc++; }
                if (cuX < W - 1) { if (numc == 2) nb2 = br; else nb3 = br; numc++; }
            }
            /* ---- stage the window around the first candidate vector (the most likely MVP) and the source block ---- */
            const int wx0 = ((la_mv_x(nb0) >> 2) - PWIN_MX) & ~3, wy0 = (la_mv_y(nb0) >> 2) - PWIN_MY;
            __syncwarp(gmask);
            fblk[2 * r] = fe.lo; fblk[2 * r + 1] = fe.hi;
            if (WIN)
            {
                /* two window rows per trip: lane r takes unit (r & 3) of row 2 k + (r >> 2) */
                const P* __restrict__ wsrc = refCU + (wy0 + half) * stride + wx0 + sub * 4;
                R4* wdst = win + half * WIN_ROW_UNITS + sub;
#pragma unroll
                for (int plane = 0; plane < 4; plane++)
                {
#pragma unroll
                    for (int k = 0; k < (WIN_H + 1) / 2; k++)
                        if (2 * k + half < WIN_H)
                            wdst[plane * WIN_PLANE_UNITS + 2 * k * WIN_ROW_UNITS] = Px<P>::load_aligned(wsrc + plane * planeSize + 2 * k * stride);
                }
            }
            __syncwarp(gmask);

            LaSearch s;
            la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);

            /* ---- CAND: SATD at each neighbour MV, no mvcost.  When every neighbour carries the same vector the first
             * candidate wins the strict-< chain whatever the costs, and its SATD is only consulted by the bidir skip rule
             * for the zero vector (slicetype.cpp:2146-2149): nothing to measure. ---- */
            const bool sameCand = numc > 0 && (numc < 2 || nb1 == nb0) && (numc < 3 || nb2 == nb0) && (numc < 4 || nb3 == nb0) && !(nb0 == 0 && bidir);
            if (sameCand)
            {
                s.mvpx = la_mv_x(nb0); s.mvpy = la_mv_y(nb0);
            }
            else if (numc)
            {
                int cc[4] = { 0, 0, 0, 0 };
                const int rounds = numc > 2 ? 2 : 1;
#pragma unroll 1
                for (int rd = 0; rd < rounds; rd++)
                {
                    const int k = 2 * rd + half;
                    const int p = la_cand_mv(s, k < numc ? k : 0);
                    const int qx = la_mv_x(p), qy = la_mv_y(p);
                    const int cost = oct_satd<P, WIN>(refCU, planeSize, stride, win, fblk, wx0, wy0, bx, by, qx, qy, win_holds_qpel(wx0, wy0, qx, qy, 0), gmask);
                    const int ca = __shfl_sync(gmask, cost, 8 * g4), cb = __shfl_sync(gmask, cost, 8 * g4 + 4);
                    if (rd == 0) { cc[0] = ca; cc[1] = cb; } else { cc[2] = ca; cc[3] = cb; }
                }
                la_upd_cand(s, cc[0], cc[1], cc[2], cc[3]);
            }
            const uint16_t* __restrict__ lutx = lut - s.mvpx;
            const uint16_t* __restrict__ luty = lut - s.mvpy;

            /* ---- START: k0 = qpel MVP (no mvcost), k1 = rounded MVP, k2 = zero ---- */
            la_enter_start(s);
            {
                const int rx = ((s.pmx + 2) >> 2) * 4, ry = ((s.pmy + 2) >> 2) * 4;
                int v[4];
                if (WIN && win_holds_qpel(wx0, wy0, s.pmx, s.pmy, 2))       /* the rounded MVP is within 2 quarter samples of it */
                {
                    v[0] = sad_row8<P>(fe, win_row_qpel<P>(win, wx0, wy0, r, s.pmx, s.pmy));
                    v[1] = sad_row8<P>(fe, win_row8<P>(win, 0, (rx >> 2) - wx0, (ry >> 2) - wy0 + r));
                }
                else
                {
                    v[0] = glob_sad_qpel<P>(refRow, planeSize, stride, fe.lo, fe.hi, s.pmx, s.pmy);
                    v[1] = glob_sad_qpel<P>(refRow, planeSize, stride, fe.lo, fe.hi, rx, ry);
                }
                if (WIN && win_holds_fpel(wx0, wy0, 0, 0, 0)) v[2] = sad_row8<P>(fe, win_row8<P>(win, 0, -wx0, -wy0 + r));
                else v[2] = glob_sad_qpel<P>(refRow, planeSize, stride, fe.lo, fe.hi, 0, 0);
                v[3] = 0;
                int tot = oct_reduce4(v, r, gmask);
                const int k = r & 3;
                const int qx = k == 1 ? rx : 0, qy = k == 1 ? ry : 0;
                if (k != 0) tot += lutx[qx] + luty[qy];
                la_upd_start(s, __shfl_sync(gmask, tot, 8 * g4), __shfl_sync(gmask, tot, 8 * g4 + 1), __shfl_sync(gmask, tot, 8 * g4 + 2));
            }

            /* ---- HEX6 + HEX3 rounds: full-pel SAD + mvcost ---- */
            {
                int tot;
                if (WIN && win_holds_fpel(wx0, wy0, s.bmx, s.bmy, 2))
                {
                    int v[8];
#pragma unroll
                    for (int k = 0; k < 6; k++)
                        v[k] = sad_row8<P>(fe, win_row8<P>(win, 0, s.bmx + la_hex2x(k + 1) - wx0, s.bmy + la_hex2y(k + 1) - wy0 + r));
                    v[6] = v[7] = 0;
                    tot = oct_reduce8(v, r, gmask);
                }
                else
                    tot = oct_pass_slow<P>(refRow, planeSize, stride, fe.lo, fe.hi, PASS_HEX6, 6, s.bmx, s.bmy, 0, r, gmask);
                const int fx = s.bmx + la_hex2x((r + 1) & 7), fy = s.bmy + la_hex2y((r + 1) & 7);
                const uint32_t key = r < 6 ? la_key(tot + lutx[fx * 4] + luty[fy * 4], r) : LA_KEY_NONE;
                bool more = la_upd_hex6(s, __reduce_min_sync(gmask, key));
                while (more)
                {
                    int t3;
                    if (WIN && win_holds_fpel(wx0, wy0, s.bmx, s.bmy, 2))
                    {
                        int v3[4];
#pragma unroll
                        for (int k = 0; k < 3; k++)
                            v3[k] = sad_row8<P>(fe, win_row8<P>(win, 0, s.bmx + la_hex2x((s.dir + k) & 7) - wx0, s.bmy + la_hex2y((s.dir + k) & 7) - wy0 + r));
                        v3[3] = 0;
                        t3 = oct_reduce4(v3, r, gmask);
                    }
                    else
                        t3 = oct_pass_slow<P>(refRow, planeSize, stride, fe.lo, fe.hi, PASS_HEX3, 3, s.bmx, s.bmy, s.dir, r, gmask);
                    const int k = r & 3;
                    const int hx = s.bmx + la_hex2x((s.dir + k) & 7), hy = s.bmy + la_hex2y((s.dir + k) & 7);
                    const uint32_t key3 = (r < 3) ? la_key(t3 + lutx[hx * 4] + luty[hy * 4], k) : LA_KEY_NONE;
                    more = la_upd_hex3(s, __reduce_min_sync(gmask, key3));
                }
            }

            /* ---- SQ8: 8-point square ---- */
            bool subpel;
            {
                int tot;
                if (WIN && win_holds_fpel(wx0, wy0, s.bmx, s.bmy, 1))
                {
                    int v[8];
#pragma unroll
                    for (int k = 0; k < 8; k++)
                        v[k] = sad_row8<P>(fe, win_row8<P>(win, 0, s.bmx + la_sq1x(k + 1) - wx0, s.bmy + la_sq1y(k + 1) - wy0 + r));
                    tot = oct_reduce8(v, r, gmask);
                }
                else
                    tot = oct_pass_slow<P>(refRow, planeSize, stride, fe.lo, fe.hi, PASS_SQ8, 8, s.bmx, s.bmy, 0, r, gmask);
                const int fx = s.bmx + la_sq1x(r + 1), fy = s.bmy + la_sq1y(r + 1);
                const uint32_t key = la_key(tot + lutx[fx * 4] + luty[fy * 4], r);
                subpel = la_upd_sq8(s, __reduce_min_sync(gmask, key), lut);
            }

            if (subpel)
            {
                /* ---- HPEL: 4 half-pel SADs ---- */
                {
                    int tot;
                    if (WIN && win_holds_qpel(wx0, wy0, s.bmx, s.bmy, 2))
                    {
                        int v[4];
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            v[k] = sad_row8<P>(fe, win_row_qpel<P>(win, wx0, wy0, r, s.bmx + la_sq1x(k + 1) * 2, s.bmy + la_sq1y(k + 1) * 2));
                        tot = oct_reduce4(v, r, gmask);
                    }
                    else
                        tot = oct_pass_slow<P>(refRow, planeSize, stride, fe.lo, fe.hi, PASS_HPEL, 4, s.bmx, s.bmy, 0, r, gmask);
                    const int k = r & 3;
                    const int qx = s.bmx + la_sq1x(k + 1) * 2, qy = s.bmy + la_sq1y(k + 1) * 2;
                    const uint32_t key = r < 4 ? la_key(tot + lutx[qx] + luty[qy], k) : LA_KEY_NONE;
                    la_upd_hpel(s, __reduce_min_sync(gmask, key));
                }
                /* ---- QPEL: SATD re-measure (k0) + 4 quarter-pel SATDs; lane k of the octet collects candidate k ---- */
                {
                    int mine = 0;
                    const bool inWin = win_holds_qpel(wx0, wy0, s.bmx, s.bmy, 1);
#pragma unroll 1
                    for (int rd = 0; rd < 3; rd++)
                    {
                        const int k = 2 * rd + half < 5 ? 2 * rd + half : 0;
                        const int cost = oct_satd<P, WIN>(refCU, planeSize, stride, win, fblk, wx0, wy0, bx, by, s.bmx + la_sq1x(k), s.bmy + la_sq1y(k), inWin, gmask);
                        const int got = __shfl_sync(gmask, cost, 8 * g4 + 4 * (r & 1));
                        if ((r >> 1) == rd) mine = got;
                    }
                    const int k = r < 5 ? r : 0;
                    const int qx = s.bmx + la_sq1x(k), qy = s.bmy + la_sq1y(k);
                    const int tot = mine + lutx[qx] + luty[qy];
                    const uint32_t key = (r >= 1 && r < 5) ? la_key(tot, r) : LA_KEY_NONE;
                    const int c0 = __shfl_sync(gmask, tot, 8 * g4);
                    la_upd_qpel(s, c0, __reduce_min_sync(gmask, key));
                }
            }
            la_finish_skip(s);

            result = la_pack_mv(s.outx, s.outy);
            if (WIN && cuX > 0)
            {
                /* the next CU's window will most likely lie around this result, 8 samples to the left: pull it towards L1
                 * now (lane r: rows r and r + 8 of two planes per half... 13 rows x 4 planes, one 32-byte sector each) */
                const P* __restrict__ nw = refCU - 8 + ((la_mv_y(result) >> 2) - PWIN_MY + r) * stride + (((la_mv_x(result) >> 2) - PWIN_MX) & ~3);
#pragma unroll
                for (int plane = 0; plane < 4; plane++)
                {
                    asm volatile("prefetch.global.L1 [%0];" :: "l"(nw + plane * planeSize));
                    if (r < WIN_H - 8) asm volatile("prefetch.global.L1 [%0];" :: "l"(nw + plane * planeSize + 8 * stride));
                }
            }
            if (r == 0)
            {
                const int cuXY = cuX + cuY * W;
                if (pubHand) pubHand[cuX] = HAND_TAG | (uint32_t)result;
                mvMirror[cuXY] = result;
                mcMirror[cuXY] = s.outcost;
                mvOut[cuXY] = result;
                mcOut[cuXY] = s.outcost;
            }
        }
        h3 = h2; h2 = h1; h1 = result;
    }
}

template <typename P>
inline size_t oct_smem_bytes(int bands, int wCU)
{
    return (size_t)bands * wCU * sizeof(unsigned long long) + (size_t)bands * 4 * (WIN_PITCH + OCT_FENC_UNITS) * sizeof(typename Px<P>::Row4);
}

#endif /* X265CU_SEARCH_OCT_CUH */
