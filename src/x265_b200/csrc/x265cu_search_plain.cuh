/* x265cu_search_plain.cuh -- the plain wavefront motion-search kernel (sm_100a).
 *
 * The search of one reference list of one estimate (encoder/slicetype.cpp:2106-2160 + the lowres branch of
 * MotionEstimate::motionEstimate, encoder/motion.cpp:571-1172) WITHOUT speculation: one warp per CU row, rows advance
 * as a wavefront, per CU the warp runs the la_core.h state machine one pass (<= 8 candidates, one per quad) at a time.
 * Least work per CU, so it is what a batch runs when the GPU is filled by independent searches (throughput-bound) or
 * when no usable hint field exists; x265cu_search.cuh holds the speculative path for latency-bound searches.
 *
 * Row groups of a few CU rows per CTA spread one slice over many SMs.  A finished CU publishes ONE 64-bit word
 * {tag = 1, packed MV}: tag and data travel in the same naturally aligned 8-byte store, so there is no separate
 * progress counter and no fence on the chain.  Inside a group the words live in shared memory, between groups in a
 * global hand-off row (L2).  A group only waits for an EARLIER work item, and items are taken from an atomic ticket
 * (take_ticket), i.e. by CTAs that are already running: the scheme cannot deadlock even when a launch does not fit the
 * GPU, whatever order the hardware starts blocks in.  Passes are branch-free: every quad always measures a
 * (valid-address) block and invalid candidates are masked out of the key reduction.
 */
#ifndef X265CU_SEARCH_PLAIN_CUH
#define X265CU_SEARCH_PLAIN_CUH

#define PLAIN_MAX_GROUP_ROWS 8
#ifndef PWIN_MX
#define PWIN_MX 2      /* window margin left of / above the first candidate's full-pel position */
#define PWIN_MY 2
#endif
/* -DX265CU_PLAIN_CLOCKS (never in the shipped library): cycles per phase of a CU step, summed over the steps of warp 0 of every CTA */
#ifdef X265CU_PLAIN_CLOCKS
__device__ unsigned long long g_plainClk[16];
__device__ unsigned long long g_plainClkB[16];   /* the same for the picture's bottom CU row (cuY == H - 1), all launches of < 700 CTAs */
__device__ unsigned long long g_plainCnt[8];    /* bottom row: steps, CAND passes, HEX3 rounds, window misses (START, HEX6, SQ8, HPEL/QPEL) */
__device__ unsigned long long g_plainTraceT[160][256];   /* [cuY][step]: globaltimer (ns) at the end of the step, launches of <= 24 CTAs */
__device__ unsigned int g_plainTraceW[160][256];         /* cycles the step spent waiting for the row below */
__device__ __forceinline__ unsigned long long gtimer_ns() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define PCLK(i) do { const long long now_ = clock64(); if (lane == 0 && gridDim.x < 700) { if (lastRow && cuY != H - 1) atomicAdd(&g_plainClk[i], (unsigned long long)(now_ - tPhase)); if (cuY == H - 1) atomicAdd(&g_plainClkB[i], (unsigned long long)(now_ - tPhase)); } tPhase = now_; } while (0)
#define PCNT(i, v) do { if (lane == 0 && gridDim.x < 700 && cuY == H - 1) atomicAdd(&g_plainCnt[i], (unsigned long long)(v)); } while (0)
#else
#define PCLK(i) do { } while (0)
#define PCNT(i, v) do { } while (0)
#endif
#ifndef PLAIN_MIN_CTAS
#define PLAIN_MIN_CTAS 3
#endif

/* WIN variant (used when a launch fills the GPU): in that regime the kernel is bound by the L1 tag stage -- a warp-wide
 * 4-byte load of 8 candidate blocks touches ~8-16 cache lines, ~900 such line look-ups per CU -- not by issue slots.
 * So the 13 x 16 x 4-plane window around the first candidate vector is staged once per CU in shared memory with
 * row-coalesced loads (56 line look-ups), and every pass whose 8 candidates lie inside it (warp-uniform test) reads
 * shared memory; a pass that leaves the window reads global memory as before. */
template <typename P, bool WIN, int RU = WIN_ROW_UNITS>
__device__ __forceinline__ void pfetch_qpel(const P* __restrict__ refLane, int planeSize, int stride, const typename Px<P>::Row4* win,
                                            int lx, int ly, int qx, int qy, typename Px<P>::Row4 out[4])
{
    if (WIN)
    {
        const int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        const int ax = lx + (qx >> 2), ay = ly + (qy >> 2), bx = lx + (qx2 >> 2), by = ly + (qy2 >> 2);
        const bool ok = (unsigned)ax <= (unsigned)(RU * 4 - 4) && (unsigned)bx <= (unsigned)(RU * 4 - 4) &&
                        (unsigned)ay <= (unsigned)(WIN_H - 4) && (unsigned)by <= (unsigned)(WIN_H - 4);
        if (__all_sync(FULL_MASK, ok)) { win_qpel<P, RU>(win, lx, ly, qx, qy, out); return; }
    }
    fetch_qpel<P>(refLane, planeSize, stride, qx, qy, out);
}

template <typename P, bool WIN, int RU = WIN_ROW_UNITS>
__device__ __forceinline__ void pfetch_fpel(const P* __restrict__ refLane, int stride, const typename Px<P>::Row4* win,
                                            int lx, int ly, int fx, int fy, typename Px<P>::Row4 out[4])
{
    if (WIN)
    {
        const bool ok = (unsigned)(lx + fx) <= (unsigned)(RU * 4 - 4) && (unsigned)(ly + fy) <= (unsigned)(WIN_H - 4);
        if (__all_sync(FULL_MASK, ok)) { win_fpel<P, RU>(win, lx, ly, fx, fy, out); return; }
    }
    fetch_off<P>(refLane, stride, fy * stride + fx, out);
}

/* ---- TMA variant (north_star: "TMA/shared-memory staging of the padded search windows").  The frame mirrors are one 3-D
 * tensor [slot x 4 planes][padded rows][stride]; a window is ONE box {WIN_W, WIN_H, 4 planes} = one cp.async.bulk.tensor.3d
 * issued by one lane, landing in the WIN layout, completion on an mbarrier.  Because the copy is asynchronous and costs the
 * warp no registers, the window of the NEXT CU is requested a whole step ahead (around this CU's first candidate vector, 8
 * samples to the left: the most likely origin) into the other of two buffers; a step whose origin was foreseen finds its
 * window in shared memory, any other requests it on demand. ---- */
/* A TMA box must start on a 16-byte boundary of its row, a window starts on a 4-sample one: the box begins at the boundary
 * below and is 16 bytes wider (8-bit: 32 samples = 8 units per row; 16-bit: 24 samples = 6 units).  The extra columns are
 * window too: fewer passes leave it. */
template <typename P> struct PlainTma;
template <> struct PlainTma<uint8_t> { enum { RU = 8, ALIGN = 16, PITCH = 448 }; };     /* 4 x 13 x 8 units + slack, 1792 bytes */
template <> struct PlainTma<uint16_t> { enum { RU = 6, ALIGN = 8, PITCH = 320 }; };     /* 4 x 13 x 6 units + slack, 2560 bytes */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* map, void* bar, int c0, int c1, int c2)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 :: "r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity)
{
    uint32_t ok;
    do
    {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    }
    while (!ok);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct alignas(64) PlainTmaMap { unsigned long long opaque[16]; };   /* CUtensorMap (128 bytes, 64-byte aligned) */

template <typename P, bool WIN, bool ONESHOT, bool TMA>
__global__ void __launch_bounds__(PLAIN_MAX_GROUP_ROWS * 32, PLAIN_MIN_CTAS)
plain_search_kernel(const JobDev* __restrict__ jobs, const SearchPlan* __restrict__ plans, const SearchItem* __restrict__ items, GeomDev g,
                    const uint16_t* __restrict__ lut, unsigned long long* gHand, unsigned int* ticket, const __grid_constant__ PlainTmaMap tmap)
{
    constexpr int RU = TMA ? (int)PlainTma<P>::RU : WIN_ROW_UNITS;     /* units of 4 samples per window row */
    extern __shared__ __align__(128) unsigned long long sHand[];  /* [blockDim / 32][W] hand-off words (+ window(s) per warp, WIN) */
    const SearchItem it = items[take_ticket(ticket)];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nRows = it.lastY - it.firstY + 1;
    const int W = g.wCU, H = g.hCU;
    typename Px<P>::Row4* win = (typename Px<P>::Row4*)(sHand + (blockDim.x >> 5) * W) + warp * WIN_PITCH;
    unsigned long long* tmaBar = NULL;
    typename Px<P>::Row4* tmaWin = NULL;
    if (TMA)
    {
        /* [warps][2] window buffers on 128-byte boundaries behind the hand-off rows, then [warps][2] barriers */
        const size_t handBytes = ((size_t)(blockDim.x >> 5) * W * sizeof(unsigned long long) + 127) & ~(size_t)127;
        typename Px<P>::Row4* base = (typename Px<P>::Row4*)((unsigned char*)sHand + handBytes);
        tmaWin = base + warp * 2 * PlainTma<P>::PITCH;
        tmaBar = (unsigned long long*)(base + (blockDim.x >> 5) * 2 * PlainTma<P>::PITCH) + warp * 2;
        if (lane == 0) { mbar_init(tmaBar, 1); mbar_init(tmaBar + 1, 1); }
        fence_proxy_async();
        win = tmaWin;
    }
    for (int i = threadIdx.x; i < nRows * W; i += blockDim.x) sHand[i] = 0;
    const SearchPlan pl = plans[it.search];
    const JobDev* __restrict__ jp = jobs + pl.job;
    const int list = pl.list;
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
    const int sq8dx = la_sq1x(q + 1), sq8dy = la_sq1y(q + 1);
    const int hpdx = la_sq1x((q + 1) & 7) * 2, hpdy = la_sq1y((q + 1) & 7) * 2;   /* quarter-pel units */
    const int qpdx = la_sq1x(q), qpdy = la_sq1y(q);

    const int rowBase = (8 * cuY + by) * stride + bx;
    /* WIN: source offset of each window unit this lane stages (unit i = lane + 32 k: plane, row, 4-sample column) */
    int wOff[(WIN_UNITS + 31) / 32];
#pragma unroll
    for (int k = 0; k < (WIN_UNITS + 31) / 32; k++)
    {
        const int i = lane + 32 * k, plane = i / WIN_PLANE_UNITS, rem = i - plane * WIN_PLANE_UNITS;
        wOff[k] = plane * planeSize + (rem / WIN_ROW_UNITS) * stride + (rem % WIN_ROW_UNITS) * 4;
    }
    int prevMv = 0;                                /* MV of (cuX + 1, cuY): our own previous result */
    int prevBl = 0, prevMb = 0;                    /* MVs of (cuX, cuY + 1) and (cuX + 1, cuY + 1) as read in the step before */
    /* TMA: the window requested ahead (buffer, tensor coordinates), barrier phases and copies in flight per buffer */
    const int tmaZ = TMA ? jp->tmaZ[list] : -1;
    bool pendValid = false;
    int pendBuf = 0, pendCol = 0, pendRow = 0;
    uint32_t phase = 0, inflight = 0;
    unsigned long long nextWord = (!lastRow && W > 1) ? below[W - 2] : 0;   /* hand-off word of this step's below-left, loaded ahead */
    typename Px<P>::Row4 fe[4], feNext[4];
#pragma unroll
    for (int y = 0; y < 4; y++)
        feNext[y] = Px<P>::load_aligned(fencPlane + rowBase + 8 * (W - 1) + y * stride);

#ifdef X265CU_PLAIN_CLOCKS
    long long tPhase = clock64();
#endif
    for (int cuX = W - 1; cuX >= 0; cuX--)
    {
        PCLK(0);
#ifdef X265CU_PLAIN_CLOCKS
        const long long tStep0 = clock64();
        long long tWaited = 0;
#endif
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
            /* the row below runs right to left: its column cuX - 1 is published last.  Only that word is new in this step
             * (below and below-right were the below-left of the two steps before), and its load was issued one step ago:
             * a row that lags the one below finds the word there, and a poll of the L2 hand-off row (~1 us round trip for
             * the bottom row of a CTA) is no longer on the chain of every step */
            int bl = 0, br = 0, mb;
            if (cuX > 0)
            {
                unsigned long long w = nextWord;
                while (!(w & HAND_TAG)) w = below[cuX - 1];
                bl = (int)(uint32_t)w;
            }
            if (cuX == W - 1) mb = hand_wait(below + cuX); else mb = prevBl;
            if (cuX < W - 1) br = prevMb;
            prevMb = mb; prevBl = bl;
            if (cuX > 1) nextWord = below[cuX - 2];
            if (numc == 0) nb0 = mb; else nb1 = mb;
            numc++;
            if (cuX > 0) { if (numc == 1) nb1 = bl; else nb2 = bl; numc++; }
            if (cuX < W - 1) { if (numc == 2) nb2 = br; else nb3 = br; numc++; }
        }
        PCLK(1);
#ifdef X265CU_PLAIN_CLOCKS
        tWaited = clock64() - tStep0;
#endif
        int lx = 0, ly = 0;
        if (WIN && TMA && tmaZ >= 0)
        {
            /* the window around the first candidate vector: foreseen one step ago, or requested now */
            const int wx0 = ((la_mv_x(nb0) >> 2) - PWIN_MX) & ~3, wy0 = (la_mv_y(nb0) >> 2) - PWIN_MY;
            const int col0 = g.marginX + 8 * cuX + wx0, row = g.marginY + 8 * cuY + wy0;
            const int col = col0 & ~(PlainTma<P>::ALIGN - 1);         /* the box starts on the 16-byte boundary below the window */
            const uint32_t winBytes = 4 * WIN_H * RU * 4 * (uint32_t)sizeof(P);
            int buf;
            if (pendValid && pendCol == col && pendRow == row)
                buf = pendBuf;
            else
            {
                buf = pendValid ? pendBuf ^ 1 : 0;
                __syncwarp();
                if (inflight & (1u << buf)) { mbar_wait(tmaBar + buf, (phase >> buf) & 1u); phase ^= 1u << buf; }
                if (lane == 0)
                {
                    fence_proxy_async();
                    mbar_expect_tx(tmaBar + buf, winBytes);
                    tma_load_3d(tmaWin + buf * PlainTma<P>::PITCH, &tmap, tmaBar + buf, col, row, tmaZ);
                }
                inflight |= 1u << buf;
            }
            mbar_wait(tmaBar + buf, (phase >> buf) & 1u);
            phase ^= 1u << buf; inflight &= ~(1u << buf);
            __syncwarp();
            win = tmaWin + buf * PlainTma<P>::PITCH;
            lx = bx - wx0 + (col0 - col); ly = by - wy0;
            pendValid = false;
            if (cuX > 0)
            {
                /* the next CU (8 samples to the left) most likely has this CU's vector as its first candidate */
                const int nb = buf ^ 1;
                if (inflight & (1u << nb)) { mbar_wait(tmaBar + nb, (phase >> nb) & 1u); phase ^= 1u << nb; inflight &= ~(1u << nb); }
                if (lane == 0)
                {
                    fence_proxy_async();
                    mbar_expect_tx(tmaBar + nb, winBytes);
                    tma_load_3d(tmaWin + nb * PlainTma<P>::PITCH, &tmap, tmaBar + nb, (col0 - 8) & ~(PlainTma<P>::ALIGN - 1), row, tmaZ);
                }
                inflight |= 1u << nb;
                pendValid = true; pendBuf = nb; pendCol = (col0 - 8) & ~(PlainTma<P>::ALIGN - 1); pendRow = row;
            }
        }
        else if (WIN)
        {
            /* stage the window around the first candidate vector (the most likely MVP) */
            const int wx0 = ((la_mv_x(nb0) >> 2) - PWIN_MX) & ~3, wy0 = (la_mv_y(nb0) >> 2) - PWIN_MY;
            const P* __restrict__ wbase = refPlane + (8 * cuY + wy0) * stride + 8 * cuX + wx0;
            __syncwarp();
            if (TMA)
            {
                /* (a weighted reference copy is not part of the tensor: same rows by loads, in the wide-row layout) */
                win = tmaWin;
                for (int i = lane; i < 4 * WIN_H * RU; i += 32)
                {
                    const int plane = i / (WIN_H * RU), rem = i - plane * (WIN_H * RU), wr = rem / RU, wc = rem % RU;
                    win[i] = Px<P>::load_aligned(wbase + plane * planeSize + wr * stride + wc * 4);
                }
            }
            else
            {
#pragma unroll
                for (int k = 0; k < (WIN_UNITS + 31) / 32; k++)
                    if (lane + 32 * k < WIN_UNITS) win[lane + 32 * k] = Px<P>::load_aligned(wbase + wOff[k]);
            }
            __syncwarp();
            lx = bx - wx0; ly = by - wy0;
        }
        PCLK(2);
        LaSearch s;
        la_search_begin(s, cuX, cuY, W, H, bidir, numc, nb0, nb1, nb2, nb3);

        /* ---- CAND: SATD at each neighbour MV, no mvcost (quads >= numc re-measure candidate 0).  When every neighbour
         * carries the same vector (3 of 4 CUs) the first candidate wins the strict-< chain whatever the costs, and its
         * SATD is only consulted by the bidir skip rule for the zero vector (slicetype.cpp:2146-2149): nothing to measure. ---- */
        const bool sameCand = numc > 0 && (numc < 2 || nb1 == nb0) && (numc < 3 || nb2 == nb0) && (numc < 4 || nb3 == nb0) && !(nb0 == 0 && bidir);
        if (sameCand)
        {
            /* la_upd_cand with equal candidates: the first one becomes the MVP, skipCost stays untouched */
            s.mvpx = la_mv_x(nb0); s.mvpy = la_mv_y(nb0);
        }
        else if (numc)
        {
            PCNT(1, 1);
            const int p = la_cand_mv(s, q < numc ? q : 0);
            typename Px<P>::Row4 r[4];
            pfetch_qpel<P, WIN, RU>(refLane, planeSize, stride, win, lx, ly, la_mv_x(p), la_mv_y(p), r);
            const int cost = quad_sum(satd4x4_abs<P>(fe, r)) >> 1;
            la_upd_cand(s, __shfl_sync(FULL_MASK, cost, 0), __shfl_sync(FULL_MASK, cost, 4),
                        __shfl_sync(FULL_MASK, cost, 8), __shfl_sync(FULL_MASK, cost, 12));
        }
        const uint16_t* __restrict__ lutx = lut - s.mvpx;
        const uint16_t* __restrict__ luty = lut - s.mvpy;

        PCLK(3);
        la_enter_start(s);
        int resume = LA_RESUME_HEX6;
        if (ONESHOT)
        {
            /* ---- one-shot: when nothing moves, every position the search visits is known once the MVP is (la_fast_path):
             * they are all measured in ONE burst of four independent rounds from the window, and the reference's decisions
             * are replayed on the costs.  A step's latency is then one pass instead of five dependent ones; a search that
             * does move (~15 % of the CUs) resumes pass by pass below. ---- */
            const int bm0x = (s.pmx + 2) >> 2, bm0y = (s.pmy + 2) >> 2;
            {
                /* the burst needs the window around the clipped MVP; it is usually where the first candidate put it */
                const int wx0 = ((s.pmx >> 2) - PWIN_MX) & ~3, wy0 = (s.pmy >> 2) - PWIN_MY;
                if (bx - wx0 != lx || by - wy0 != ly)
                {
                    const P* __restrict__ wbase = refPlane + (8 * cuY + wy0) * stride + 8 * cuX + wx0;
                    __syncwarp();
#pragma unroll
                    for (int k = 0; k < (WIN_UNITS + 31) / 32; k++)
                        if (lane + 32 * k < WIN_UNITS) win[lane + 32 * k] = Px<P>::load_aligned(wbase + wOff[k]);
                    __syncwarp();
                    lx = bx - wx0; ly = by - wy0;
                }
            }
            /* round 1: q0 qpel MVP (no mvcost), q1 rounded MVP, q3..6 the 4 half-pel points around pm */
            int cost1;
            {
                const int hq = (q >= 3 && q < 7) ? q - 2 : 0;
                const int qx = (q == 1 ? bm0x * 4 : s.pmx) + la_sq1x(hq) * 2;
                const int qy = (q == 1 ? bm0y * 4 : s.pmy) + la_sq1y(hq) * 2;
                typename Px<P>::Row4 r[4];
                win_qpel<P>(win, lx, ly, qx, qy, r);
                const int mvc = q == 0 ? 0 : lutx[qx] + luty[qy];
                cost1 = quad_sum(sad4x4<P>(fe, r)) + mvc;
            }
            /* rounds 2, 3: hexagon (6) + square points 1, 2 | square points 3..8 + the zero MV (from global memory) */
            int cost2, cost3;
            {
                const int r2x = bm0x + (q < 6 ? la_hex2x(q + 1) : la_sq1x(q - 5)), r2y = bm0y + (q < 6 ? la_hex2y(q + 1) : la_sq1y(q - 5));
                const int r3x = q < 6 ? bm0x + la_sq1x(q + 3) : 0, r3y = q < 6 ? bm0y + la_sq1y(q + 3) : 0;
                typename Px<P>::Row4 r2[4], r3[4];
                win_fpel<P>(win, lx, ly, r2x, r2y, r2);
                if (q < 6) win_fpel<P>(win, lx, ly, r3x, r3y, r3);
                else fetch_off<P>(refLane, stride, 0, r3);
                cost2 = quad_sum(sad4x4<P>(fe, r2)) + lutx[r2x * 4] + luty[r2y * 4];
                cost3 = quad_sum(sad4x4<P>(fe, r3)) + lutx[r3x * 4] + luty[r3y * 4];
            }
            /* round 4: SATD at pm and the 4 quarter-pel points around it */
            int cost4;
            {
                const int k = q < 5 ? q : 0;
                const int qx = s.pmx + la_sq1x(k), qy = s.pmy + la_sq1y(k);
                typename Px<P>::Row4 r[4];
                win_qpel<P>(win, lx, ly, qx, qy, r);
                cost4 = (quad_sum(satd4x4_abs<P>(fe, r)) >> 1) + lutx[qx] + luty[qy];
            }
            const int c0 = __shfl_sync(FULL_MASK, cost1, 0), c1 = __shfl_sync(FULL_MASK, cost1, 4), c2 = __shfl_sync(FULL_MASK, cost3, 24);
            const uint32_t hpelKey = warp_min_key(q >= 3 && q < 7, cost1, q - 3);
            const uint32_t hexKey = warp_min_key(q < 6, cost2, q);
            const uint32_t sqKey = __reduce_min_sync(FULL_MASK, q < 6 ? la_key(cost3, q + 2) : la_key(cost2, q - 6));
            const int qc0 = __shfl_sync(FULL_MASK, cost4, 0);
            const uint32_t qpelKey = warp_min_key(q >= 1 && q < 5, cost4, q);
            resume = la_fast_path(s, c0, c1, c2, hexKey, sqKey, hpelKey, qc0, qpelKey, lut);
        }
        else
        {
            /* ---- START: q0 = qpel MVP (no mvcost), q1 = rounded MVP, q2 = zero ---- */
            const int qx = q == 0 ? s.pmx : (q == 1 ? ((s.pmx + 2) >> 2) * 4 : 0);
            const int qy = q == 0 ? s.pmy : (q == 1 ? ((s.pmy + 2) >> 2) * 4 : 0);
            typename Px<P>::Row4 r[4];
            pfetch_qpel<P, WIN, RU>(refLane, planeSize, stride, win, lx, ly, qx, qy, r);
            const int mvc = q == 0 ? 0 : lutx[qx] + luty[qy];
            const int cost = quad_sum(sad4x4<P>(fe, r)) + mvc;
            la_upd_start(s, __shfl_sync(FULL_MASK, cost, 0), __shfl_sync(FULL_MASK, cost, 4), __shfl_sync(FULL_MASK, cost, 8));
        }
        PCLK(4);

        if (resume != LA_RESUME_DONE)
        {
            /* ---- HEX6 + HEX3 rounds: full-pel SAD + mvcost ---- */
            bool more = resume == LA_RESUME_HEX3;
            if (resume == LA_RESUME_HEX6)
            {
                const int fx = s.bmx + hex6dx, fy = s.bmy + hex6dy;
                typename Px<P>::Row4 r[4];
                pfetch_fpel<P, WIN, RU>(refLane, stride, win, lx, ly, fx, fy, r);
                const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
                more = la_upd_hex6(s, warp_min_key(q < 6, cost, q));
            }
            while (more)
            {
                PCNT(2, 1);
                const int hdx = la_hex2x((s.dir + q) & 7), hdy = la_hex2y((s.dir + q) & 7);
                const int hx = s.bmx + hdx, hy = s.bmy + hdy;
                typename Px<P>::Row4 r3[4];
                pfetch_fpel<P, WIN, RU>(refLane, stride, win, lx, ly, hx, hy, r3);
                const int c3 = quad_sum(sad4x4<P>(fe, r3)) + lutx[hx * 4] + luty[hy * 4];
                more = la_upd_hex3(s, warp_min_key(q < 3, c3, q));
            }
            PCLK(5);

            /* ---- SQ8: 8-point square ---- */
            bool subpel = true;
            if (resume <= LA_RESUME_SQ8)
            {
                const int fx = s.bmx + sq8dx, fy = s.bmy + sq8dy;
                typename Px<P>::Row4 r[4];
                pfetch_fpel<P, WIN, RU>(refLane, stride, win, lx, ly, fx, fy, r);
                const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[fx * 4] + luty[fy * 4];
                subpel = la_upd_sq8(s, warp_min_key(true, cost, q), lut);
            }
            PCLK(6);
            if (subpel)
            {
                /* ---- HPEL: 4 half-pel SADs ---- */
                if (resume <= LA_RESUME_HPEL)
                {
                    const int qx = s.bmx + hpdx, qy = s.bmy + hpdy;
                    typename Px<P>::Row4 r[4];
                    pfetch_qpel<P, WIN, RU>(refLane, planeSize, stride, win, lx, ly, qx, qy, r);
                    const int cost = quad_sum(sad4x4<P>(fe, r)) + lutx[qx] + luty[qy];
                    la_upd_hpel(s, warp_min_key(q < 4, cost, q));
                }
                PCLK(7);
                /* ---- QPEL: SATD re-measure (q0) + 4 quarter-pel SATDs ---- */
                {
                    const int qx = s.bmx + qpdx, qy = s.bmy + qpdy;
                    typename Px<P>::Row4 r[4];
                    pfetch_qpel<P, WIN, RU>(refLane, planeSize, stride, win, lx, ly, qx, qy, r);
                    const int cost = (quad_sum(satd4x4_abs<P>(fe, r)) >> 1) + lutx[qx] + luty[qy];
                    const int c0 = __shfl_sync(FULL_MASK, cost, 0);
                    la_upd_qpel(s, c0, warp_min_key(q >= 1 && q < 5, cost, q));
                }
            }
        }
        PCLK(8);
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
        PCLK(9);
        if (TMA && cuX == 0 && inflight)
        {
            /* no copy may still be writing into this CTA's shared memory when it retires */
            for (int b2 = 0; b2 < 2; b2++)
                if (inflight & (1u << b2)) mbar_wait(tmaBar + b2, (phase >> b2) & 1u);
        }
#ifdef X265CU_PLAIN_CLOCKS
        if (lane == 0 && gridDim.x < 700 && lastRow && cuY != H - 1) atomicAdd(&g_plainClk[15], 1ull);
        if (lane == 0 && gridDim.x < 700 && cuY == H - 1) atomicAdd(&g_plainClkB[15], 1ull);
        if (lane == 0 && gridDim.x <= 24 && cuY < 160 && W <= 256) { g_plainTraceT[cuY][W - 1 - cuX] = gtimer_ns(); g_plainTraceW[cuY][W - 1 - cuX] = (unsigned int)tWaited; }
#endif
    }
}


#endif /* X265CU_SEARCH_PLAIN_CUH */
