/* x265cu_kernels.cuh -- the CUDA kernels of libx265cu.so (hand-written for sm_100a).
 *
 * Kernel            replaces (x265_1.9/source)                                  bound
 * ----------------  ----------------------------------------------------------  ----------------
 * lowres_init       frame_init_lowres_core pixel.cpp:549 + extendPicBorder :908   HBM (8*Np*P bytes)
 * intra_estimate    lowresIntraEstimate slicetype.cpp:230-336, intrapred.cpp      int ALU
 * search            estimateCUCost slicetype.cpp:2068 + motionEstimate (lowres)   dependency latency
 *                   motion.cpp:571; wavefront over CUs, one CTA per (job, slice)  / int ALU
 * cost              estimateCUCost without searches (cost-only estimates)         L2/int ALU
 * weight_planes     weight_pp_c pixel.cpp:463 over the 4 padded planes            HBM
 * weight_cost       weightCostLuma slicetype.cpp:338-371                          HBM
 * pixelcmp_*        sad/satd/sa8d primitives pixel.cpp:39-322                     HBM
 * frame_var         pixel_var<16>/<8> pixel.cpp:649 (acEnergyCu)                  HBM
 *
 * Tensor cores are not used anywhere: this is byte/integer work (SAD, Hadamard, compares).
 */
#ifndef X265CU_KERNELS_CUH
#define X265CU_KERNELS_CUH

#include "x265cu_dev.cuh"

#define X265CU_SA8D_16x16_K 3

/* one result array of a batch on its way from the device staging area into the caller's (mapped, pinned) host array */
struct ScatterDev { const uint8_t* src; uint8_t* dst; unsigned bytes; unsigned srcOff; };

/* CTA per array: 16-byte coalesced stores over PCIe (both sides are 16-byte aligned: staging records and the host
 * layer's arrays are), byte tail */
__global__ void __launch_bounds__(256) scatter_results_kernel(const ScatterDev* __restrict__ e)
{
    const ScatterDev s = e[blockIdx.x];
    unsigned done = 0;
    if ((((size_t)s.src | (size_t)s.dst) & 15) == 0)
    {
        const unsigned n16 = s.bytes >> 4;
        const uint4* a = (const uint4*)s.src;
        uint4* b = (uint4*)s.dst;
        for (unsigned i = threadIdx.x; i < n16; i += blockDim.x) b[i] = a[i];
        done = n16 << 4;
    }
    for (unsigned i = done + threadIdx.x; i < s.bytes; i += blockDim.x) s.dst[i] = s.src[i];
}

/* the same with the entries passed by value (no argument upload): the intra outputs of a few frames */
#define SCATTER_SMALL 32
struct ScatterBatch { ScatterDev e[SCATTER_SMALL]; };
__global__ void __launch_bounds__(256) scatter_small_kernel(ScatterBatch b)
{
    const ScatterDev s = b.e[blockIdx.x];
    unsigned done = 0;
    if ((((size_t)s.src | (size_t)s.dst) & 15) == 0)
    {
        const unsigned n16 = s.bytes >> 4;
        const uint4* a = (const uint4*)s.src;
        uint4* d = (uint4*)s.dst;
        for (unsigned i = threadIdx.x; i < n16; i += blockDim.x) d[i] = a[i];
        done = n16 << 4;
    }
    for (unsigned i = done + threadIdx.x; i < s.bytes; i += blockDim.x) s.dst[i] = s.src[i];
}

struct GeomDev
{
    int width, lines, stride, marginX, marginY, paddedLines;
    int wCU, hCU, nCU;
    int64_t planeSize, padOffset;
};

/* ===========================================================================================
 * lowres_init: one thread = 16 bytes of one padded output row of each of the four planes.
 * Rows/columns in the margins recompute the clamped source position, so the border replication
 * of extendPicBorder is fused into the same pass: no second kernel, no inter-block dependency.
 * FILTER(a,b,c,d) = avg(avg(a,b), avg(c,d)) with round-up at both levels == chained __vavgu4.
 * =========================================================================================== */
template <typename P>
__device__ __forceinline__ int lowres_px(const P* r0, const P* r1, int c0, int c1)
{
    int a = ((int)r0[c0] + (int)r1[c0] + 1) >> 1;
    int b = ((int)r0[c1] + (int)r1[c1] + 1) >> 1;
    return (a + b + 1) >> 1;
}

/* frames of one pre-lookahead list share a launch: blockIdx.z (lowres) / blockIdx.y (variance, intra) = frame */
#define PRE_BATCH 8
struct LowresBatch { const void* src[PRE_BATCH]; void* planes[PRE_BATCH]; int64_t pitch[PRE_BATCH]; };
struct VarBatch { const void* y[PRE_BATCH]; const void* u[PRE_BATCH]; const void* v[PRE_BATCH]; unsigned int* energy[PRE_BATCH]; unsigned long long* sums[PRE_BATCH];
                  int64_t ys[PRE_BATCH], cs[PRE_BATCH];
                  void* uKeep[PRE_BATCH]; void* vKeep[PRE_BATCH]; };      /* compact copies of the chroma planes kept per frame slot (x265cu_wp.cuh), or NULL */

/* One thread = LOWRES_UNIT(P) = 16 / sizeof(P) consecutive samples of one padded output row of all four planes: one 16-byte
 * store per plane.  Threads are numbered over the units of the whole padded plane (unitsPerRow x paddedLines), so every
 * thread of the grid has work.  Interior units whose source rows are 16-byte aligned read 2 x 16 bytes + 1 sample per
 * source row and compute with packed rounded averages; units that lie entirely in the left / right margin replicate one
 * value; the rest (mixed units, unaligned sources) go sample by sample. */
#define LOWRES_UNIT(P) (16 / (int)sizeof(P))

/* 8 packed source samples lo|hi of a vertical average + the 9th -> 4 outputs at even positions and 4 at odd + 1/2 */
__device__ __forceinline__ void lowres_pack8(uint32_t lo, uint32_t hi, uint32_t next, uint32_t& full, uint32_t& half)
{
    const uint32_t e = __byte_perm(lo, hi, 0x6420), o = __byte_perm(lo, hi, 0x7531);
    const uint32_t f = (e >> 8) | (next << 24);
    full = __vavgu4(e, o);
    half = __vavgu4(o, f);
}
/* the same for 16-bit samples: 4 packed samples lo|hi + the 5th -> 2 + 2 outputs */
__device__ __forceinline__ void lowres_pack16(uint32_t lo, uint32_t hi, uint32_t next, uint32_t& full, uint32_t& half)
{
    const uint32_t e = __byte_perm(lo, hi, 0x5410), o = __byte_perm(lo, hi, 0x7632);
    const uint32_t f = (e >> 16) | (next << 16);
    full = __vavgu2(e, o);
    half = __vavgu2(o, f);
}

template <typename P>
__device__ __forceinline__ void lowres_init_body(const P* __restrict__ src, int64_t srcPitch, P* __restrict__ planes, const GeomDev& g)
{
    const int V = LOWRES_UNIT(P);
    const int padW = g.width + 2 * g.marginX;
    const int unitsPerRow = (padW + V - 1) / V;
    const int unit = blockIdx.x * blockDim.x + threadIdx.x;
    if (unit >= unitsPerRow * g.paddedLines) return;
    const int oy = unit / unitsPerRow;                          /* padded row */
    const int ox0 = (unit - oy * unitsPerRow) * V;              /* padded column of the first sample */
    int yy = oy - g.marginY;
    yy = yy < 0 ? 0 : (yy > g.lines - 1 ? g.lines - 1 : yy);
    const P* r0 = src + (int64_t)(2 * yy) * srcPitch;
    const P* r1 = r0 + srcPitch;
    const P* r2 = r1 + srcPitch;
    P* d0 = planes + (int64_t)oy * g.stride + ox0;
    P* dh = d0 + g.planeSize;
    P* dv = dh + g.planeSize;
    P* dc = dv + g.planeSize;
    const int x0 = ox0 - g.marginX;
    const bool whole = ox0 + V <= padW && (((uintptr_t)planes | (uintptr_t)(g.stride * (int)sizeof(P)) | (uintptr_t)(g.planeSize * (int64_t)sizeof(P))) & 15) == 0;
    if (whole && x0 >= 0 && x0 + V <= g.width && (((uintptr_t)src | (uintptr_t)(srcPitch * (int64_t)sizeof(P))) & 15) == 0 && ((2 * x0 * (int)sizeof(P)) & 15) == 0)
    {
        /* interior fast path: 2 V + 1 source samples per row */
        const uint4* s0 = (const uint4*)(r0 + 2 * x0);
        const uint4* s1 = (const uint4*)(r1 + 2 * x0);
        const uint4* s2 = (const uint4*)(r2 + 2 * x0);
        const uint4 a0 = __ldg(s0), a1 = __ldg(s0 + 1), b0 = __ldg(s1), b1 = __ldg(s1 + 1), c0 = __ldg(s2), c1 = __ldg(s2 + 1);
        const uint32_t ax = __ldg(r0 + 2 * x0 + 2 * V), bx = __ldg(r1 + 2 * x0 + 2 * V), cx = __ldg(r2 + 2 * x0 + 2 * V);
        uint32_t v01[9], v12[9];
        if (sizeof(P) == 1)
        {
            v01[0] = __vavgu4(a0.x, b0.x); v01[1] = __vavgu4(a0.y, b0.y); v01[2] = __vavgu4(a0.z, b0.z); v01[3] = __vavgu4(a0.w, b0.w);
            v01[4] = __vavgu4(a1.x, b1.x); v01[5] = __vavgu4(a1.y, b1.y); v01[6] = __vavgu4(a1.z, b1.z); v01[7] = __vavgu4(a1.w, b1.w);
            v12[0] = __vavgu4(b0.x, c0.x); v12[1] = __vavgu4(b0.y, c0.y); v12[2] = __vavgu4(b0.z, c0.z); v12[3] = __vavgu4(b0.w, c0.w);
            v12[4] = __vavgu4(b1.x, c1.x); v12[5] = __vavgu4(b1.y, c1.y); v12[6] = __vavgu4(b1.z, c1.z); v12[7] = __vavgu4(b1.w, c1.w);
        }
        else
        {
            v01[0] = __vavgu2(a0.x, b0.x); v01[1] = __vavgu2(a0.y, b0.y); v01[2] = __vavgu2(a0.z, b0.z); v01[3] = __vavgu2(a0.w, b0.w);
            v01[4] = __vavgu2(a1.x, b1.x); v01[5] = __vavgu2(a1.y, b1.y); v01[6] = __vavgu2(a1.z, b1.z); v01[7] = __vavgu2(a1.w, b1.w);
            v12[0] = __vavgu2(b0.x, c0.x); v12[1] = __vavgu2(b0.y, c0.y); v12[2] = __vavgu2(b0.z, c0.z); v12[3] = __vavgu2(b0.w, c0.w);
            v12[4] = __vavgu2(b1.x, c1.x); v12[5] = __vavgu2(b1.y, c1.y); v12[6] = __vavgu2(b1.z, c1.z); v12[7] = __vavgu2(b1.w, c1.w);
        }
        v01[8] = (ax + bx + 1) >> 1;
        v12[8] = (bx + cx + 1) >> 1;
        uint32_t o0[4], oh[4], ov[4], oc[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
            const uint32_t mask = sizeof(P) == 1 ? 0xffu : 0xffffu;
            if (sizeof(P) == 1)
            {
                lowres_pack8(v01[2 * j], v01[2 * j + 1], v01[2 * j + 2] & mask, o0[j], oh[j]);
                lowres_pack8(v12[2 * j], v12[2 * j + 1], v12[2 * j + 2] & mask, ov[j], oc[j]);
            }
            else
            {
                lowres_pack16(v01[2 * j], v01[2 * j + 1], v01[2 * j + 2] & mask, o0[j], oh[j]);
                lowres_pack16(v12[2 * j], v12[2 * j + 1], v12[2 * j + 2] & mask, ov[j], oc[j]);
            }
        }
        *(uint4*)d0 = make_uint4(o0[0], o0[1], o0[2], o0[3]);
        *(uint4*)dh = make_uint4(oh[0], oh[1], oh[2], oh[3]);
        *(uint4*)dv = make_uint4(ov[0], ov[1], ov[2], ov[3]);
        *(uint4*)dc = make_uint4(oc[0], oc[1], oc[2], oc[3]);
        return;
    }
    if (whole && (x0 + V <= 0 || x0 >= g.width))
    {
        /* entirely in the left / right margin: the edge sample of each plane, replicated */
        const int xx = x0 < 0 ? 0 : g.width - 1;
        const int c0 = 2 * xx, c1 = c0 + 1, c2 = c0 + 2;
        uint32_t e0 = (uint32_t)lowres_px(r0, r1, c0, c1), eh = (uint32_t)lowres_px(r0, r1, c1, c2);
        uint32_t ev = (uint32_t)lowres_px(r1, r2, c0, c1), ec = (uint32_t)lowres_px(r1, r2, c1, c2);
        const uint32_t rep = sizeof(P) == 1 ? 0x01010101u : 0x00010001u;
        e0 *= rep; eh *= rep; ev *= rep; ec *= rep;
        *(uint4*)d0 = make_uint4(e0, e0, e0, e0);
        *(uint4*)dh = make_uint4(eh, eh, eh, eh);
        *(uint4*)dv = make_uint4(ev, ev, ev, ev);
        *(uint4*)dc = make_uint4(ec, ec, ec, ec);
        return;
    }
    for (int i = 0; i < V && ox0 + i < padW; i++)
    {
        int xx = x0 + i;
        xx = xx < 0 ? 0 : (xx > g.width - 1 ? g.width - 1 : xx);
        const int c0 = 2 * xx, c1 = c0 + 1, c2 = c0 + 2;
        d0[i] = (P)lowres_px(r0, r1, c0, c1);
        dh[i] = (P)lowres_px(r0, r1, c1, c2);
        dv[i] = (P)lowres_px(r1, r2, c0, c1);
        dc[i] = (P)lowres_px(r1, r2, c1, c2);
    }
}

template <typename P>
__global__ void __launch_bounds__(256) lowres_init_kernel(const P* __restrict__ src, int64_t srcPitch, P* __restrict__ planes, GeomDev g)
{
    lowres_init_body<P>(src, srcPitch, planes, g);
}

template <typename P>
__global__ void __launch_bounds__(256) lowres_init_batch_kernel(LowresBatch b, GeomDev g)
{
    lowres_init_body<P>((const P*)b.src[blockIdx.z], b.pitch[blockIdx.z], (P*)b.planes[blockIdx.z], g);
}

#include "x265cu_intra.cuh"

/* ===========================================================================================
 * search / cost kernels
 * =========================================================================================== */
struct JobDev
{
    const void* fenc;          /* sample (0,0) of plane 0 of frames[b] */
    const void* ref0w;         /* L0 reference for the SEARCH (weighted copy when isWeighted) */
    const void* ref0;          /* un-weighted L0 reference (bidir MC, co-located) */
    const void* ref1;
    int* mvs[2];               /* device mirrors of lowresMvs[l][d-1] (packed int16 x | y << 16) */
    int* mvCosts[2];
    uint16_t* lowresCosts;     /* mirror of lowresCosts[d0][d1] */
    int* rowSatds;             /* mirror of rowSatds[d0][d1] */
    const int* intraCost;
    const int* invQ;           /* or NULL */
    /* packed output record of this job (device staging, copied to the host in one transfer) */
    unsigned long long* outSums;   /* costEst, costEstAq, intraMbs */
    int* outRows;
    uint16_t* outLowresCosts;
    int* outMvs[2];
    int* outMvCosts[2];
    int d0, d1;
    int doSearch[2];
    int bidir;
    int tmaZ[2];               /* first plane of the list's search reference in the frame-mirror tensor (TMA), -1: not in it */
};

#include "x265cu_search.cuh"
#include "x265cu_search_plain.cuh"
#include "x265cu_search_oct.cuh"
#include "x265cu_wp.cuh"
#include "x265cu_me.cuh"

/* cost-only estimates (both bDoSearch false): every CU is independent.  grid = (hCU, jobs);
 * a block owns one CU row of one job, so rowSatds needs no atomics. */
template <typename P>
__global__ void __launch_bounds__(128) cost_kernel(const JobDev* __restrict__ jobs, const int* __restrict__ jobIdx, GeomDev g)
{
    __shared__ int sRow;
    __shared__ unsigned long long sSums[3];
    const JobDev& job = jobs[jobIdx[blockIdx.y]];
    const int cuY = blockIdx.x;
    const int W = g.wCU, H = g.hCU;
    if (threadIdx.x == 0) sRow = 0;
    if (threadIdx.x < 3) sSums[threadIdx.x] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int bidir = job.bidir;
    const int hasQ = job.invQ != NULL;
    int rowSum = 0, accIntra = 0;
    long long accCost = 0, accAq = 0;
    if (!bidir)
    {
        for (int cuX = threadIdx.x; cuX < W; cuX += blockDim.x)
        {
            const int cuXY = cuX + cuY * W;
            LaCuResult res = la_cu_finish(cuX, cuY, W, H, 0, job.mvCosts[0][cuXY], LA_COST_MAX, LA_COST_MAX, LA_COST_MAX,
                                          job.intraCost[cuXY], hasQ, hasQ ? job.invQ[cuXY] : 256);
            if (res.scored) { accCost += res.bcost; accAq += res.bcostAq; accIntra += res.intraMb; }
            rowSum += res.bcostAq;
            job.lowresCosts[cuXY] = res.lowresCost;
            job.outLowresCosts[cuXY] = res.lowresCost;
        }
    }
    else
    {
        const P* fencPlane = (const P*)job.fenc;
        RefPlanes<P> r0 = { (const P*)job.ref0, g.planeSize, g.stride };
        RefPlanes<P> r1 = { (const P*)job.ref1, g.planeSize, g.stride };
        const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
        const int nWarps = blockDim.x >> 5;
        /* a warp takes 8 adjacent CUs (one per quad) and measures both bidir candidates of each: the average of the two
         * motion-compensated blocks, then the average of the two co-located blocks (aligned rows: one load per row) */
        /* the vectors of the NEXT round are requested before this round's pixels (the stores below keep the compiler from
         * hoisting them): the vector -> pixel address dependency is off the critical path of every round but the first */
        int mv0n = 0, mv1n = 0;
        if (warp * 8 + q < W) { mv0n = job.mvs[0][warp * 8 + q + cuY * W]; mv1n = job.mvs[1][warp * 8 + q + cuY * W]; }
        for (int base = warp * 8; base < W; base += nWarps * 8)
        {
            const int cuX = base + q;
            const int valid = cuX < W;
            const int cuXY = (valid ? cuX : 0) + cuY * W;
            const int mv0 = mv0n, mv1 = mv1n;
            {
                const int nx = cuX + nWarps * 8;
                if (nx < W) { mv0n = job.mvs[0][nx + cuY * W]; mv1n = job.mvs[1][nx + cuY * W]; }
            }
            /* what la_cu_finish reads, requested together with the pixels */
            int mc0 = 0, mc1 = 0, icost = 0, invq = 256;
            if (valid && sub == 0)
            {
                mc0 = job.mvCosts[0][cuXY]; mc1 = job.mvCosts[1][cuXY]; icost = job.intraCost[cuXY];
                if (hasQ) invq = job.invQ[cuXY];
            }
            int partMc = 0, partCo = 0;
            if (valid)
            {
                const int px = 8 * cuX + bx, py = 8 * cuY + by;
                typename Px<P>::Row4 fe[4], a[4], b[4];
#pragma unroll
                for (int y = 0; y < 4; y++)
                    fe[y] = Px<P>::load_aligned(fencPlane + (int64_t)(py + y) * g.stride + px);
                mc_fetch4x4<P>(r0, px, py, la_mv_x(mv0), la_mv_y(mv0), a);
                mc_fetch4x4<P>(r1, px, py, la_mv_x(mv1), la_mv_y(mv1), b);
#pragma unroll
                for (int y = 0; y < 4; y++) a[y] = Px<P>::avg(a[y], b[y]);
                partMc = satd4x4_abs<P>(fe, a);
#pragma unroll
                for (int y = 0; y < 4; y++)
                    a[y] = Px<P>::avg(Px<P>::load_aligned(r0.p0 + (int64_t)(py + y) * g.stride + px),
                                      Px<P>::load_aligned(r1.p0 + (int64_t)(py + y) * g.stride + px));
                partCo = satd4x4_abs<P>(fe, a);
            }
            const int cost = quad_sum(partMc) >> 1, other = quad_sum(partCo) >> 1;
            if (valid && sub == 0)
            {
                LaCuResult res = la_cu_finish(cuX, cuY, W, H, 1, mc0, mc1, cost, other, icost, hasQ, invq);
                if (res.scored) { accCost += res.bcost; accAq += res.bcostAq; }
                rowSum += res.bcostAq;
                job.lowresCosts[cuXY] = res.lowresCost;
                job.outLowresCosts[cuXY] = res.lowresCost;
            }
        }
    }
    rowSum = warp_sum(rowSum);
    accIntra = warp_sum(accIntra);
    /* 64-bit sums: reduce as two 32-bit halves is unnecessary -- per-row totals fit 32 bits */
    int c32 = warp_sum((int)accCost), a32 = warp_sum((int)accAq);
    if (lane == 0)
    {
        atomicAdd(&sRow, rowSum);
        atomicAdd(&sSums[0], (unsigned long long)(unsigned int)c32);
        atomicAdd(&sSums[1], (unsigned long long)(unsigned int)a32);
        atomicAdd(&sSums[2], (unsigned long long)(unsigned int)accIntra);
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        job.rowSatds[cuY] = sRow;
        job.outRows[cuY] = sRow;
    }
    if (threadIdx.x < 3)
        atomicAdd(&job.outSums[threadIdx.x], sSums[threadIdx.x]);
}

/* ===========================================================================================
 * weighted prediction helpers
 * =========================================================================================== */
struct WeightDev
{
    const void* src;     /* buffer[0] of the reference (start of the padded plane 0) */
    void* dst;           /* weighted copy, same layout (4 padded planes) */
    int scale, round, shift, offset;   /* weight_pp arguments (round/shift already include the 14-depth correction) */
};

/* weight_pp_c (pixel.cpp:463-488) over all four padded planes: dst = clip(((w0 * (src << corr) + round) >> shift) + offset) */
template <typename P>
__global__ void __launch_bounds__(256) weight_planes_kernel(const WeightDev* __restrict__ items, int64_t totalSamples, int correction, int pixelMax)
{
    const WeightDev w = items[blockIdx.y];
    const P* src = (const P*)w.src;
    P* dst = (P*)w.dst;
    const int64_t per = 16 / sizeof(P);
    for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * per; i < totalSamples; i += (int64_t)gridDim.x * blockDim.x * per)
    {
        uint4 v = __ldg((const uint4*)(src + i));
        P* pv = (P*)&v;
#pragma unroll
        for (int k = 0; k < (int)per; k++)
        {
            int val = (int)(int16_t)((int)pv[k] << correction);
            int r = ((w.scale * val + w.round) >> w.shift) + w.offset;
            pv[k] = (P)(r < 0 ? 0 : (r > pixelMax ? pixelMax : r));
        }
        *(uint4*)(dst + i) = v;
    }
}

struct WeightCostDev
{
    const void* fenc;    /* sample (0,0) of plane 0 */
    const void* ref;     /* sample (0,0) of plane 0 of the reference */
    const int* intraCost;
    int weighted, scale, round, shift, offset;
};

/* weightCostLuma (slicetype.cpp:338-371): the weighted reference is never materialised, the weight
 * is applied to the reference samples on the fly.  grid = (ceil(nCU / 8 / warpsPerBlock), items). */
template <typename P>
__global__ void __launch_bounds__(256) weight_cost_kernel(const WeightCostDev* __restrict__ items, unsigned int* __restrict__ costs,
                                                           GeomDev g, int correction, int pixelMax)
{
    const WeightCostDev w = items[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const P* fenc = (const P*)w.fenc;
    const P* ref = (const P*)w.ref;
    unsigned int acc = 0;
    for (int base = (blockIdx.x * nWarps + warp) * 8; base < g.nCU; base += gridDim.x * nWarps * 8)
    {
        const int mb = base + q;
        const int valid = mb < g.nCU;
        int part = 0;
        if (valid)
        {
            const int cuX = mb % g.wCU, cuY = mb / g.wCU;
            const int64_t off = (int64_t)(8 * cuY + by) * g.stride + 8 * cuX + bx;
            int d[4][4];
#pragma unroll
            for (int y = 0; y < 4; y++)
            {
                int a[4], b[4];
                Px<P>::unpack(Px<P>::load_aligned(ref + off + (int64_t)y * g.stride), a);
                Px<P>::unpack(Px<P>::load_aligned(fenc + off + (int64_t)y * g.stride), b);
#pragma unroll
                for (int x = 0; x < 4; x++)
                {
                    int v = a[x];
                    if (w.weighted)
                    {
                        int val = (int)(int16_t)(v << correction);
                        v = ((w.scale * val + w.round) >> w.shift) + w.offset;
                        v = v < 0 ? 0 : (v > pixelMax ? pixelMax : v);
                    }
                    d[y][x] = v - b[x];
                }
            }
            part = hadamard4x4_abs(d);
        }
        int satd = quad_sum(part) >> 1;
        if (valid && sub == 0)
        {
            int ic = w.intraCost[mb];
            acc += (unsigned int)(satd < ic ? satd : ic);
        }
    }
    acc = (unsigned int)warp_sum((int)acc);
    if (lane == 0 && acc)
        atomicAdd(&costs[blockIdx.y], acc);
}

/* ===========================================================================================
 * pixel primitives as batch kernels (common/pixel.cpp:39-322)
 * =========================================================================================== */
/* 8-point Hadamard abs-sum pieces for sa8d: lane owns a 4x4 quadrant of the 8x8 difference block;
 * the 8x8 transform = 4x4 transform inside each quadrant, then a 2x2 butterfly across quadrants. */
__device__ __forceinline__ int sa8d_quad_abs(int d[4][4], int sub)
{
    /* in-quadrant 4x4 Hadamard (no abs) */
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1];
        int s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
        d[y][0] = s01 + s23; d[y][1] = d01 + d23; d[y][2] = s01 - s23; d[y][3] = d01 - d23;
    }
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        int s01 = d[0][x] + d[1][x], d01 = d[0][x] - d[1][x];
        int s23 = d[2][x] + d[3][x], d23 = d[2][x] - d[3][x];
        d[0][x] = s01 + s23; d[1][x] = d01 + d23; d[2][x] = s01 - s23; d[3][x] = d01 - d23;
    }
    int sum = 0;
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
        for (int x = 0; x < 4; x++)
        {
            int v = d[y][x];
            int h = __shfl_xor_sync(FULL_MASK, v, 1);          /* horizontal neighbour quadrant */
            v = (sub & 1) ? h - v : v + h;
            int w = __shfl_xor_sync(FULL_MASK, v, 2);          /* vertical neighbour quadrant */
            v = (sub & 2) ? w - v : v + w;
            sum += abs(v);
        }
    return sum;
}

/* pu[partitionFromSizes(w, h)].sad / .satd for any of the 25 luma PU shapes (common/pixel.cpp:954-1004): n
 * block pairs of one shape at arbitrary sample offsets, a warp per pair, the lanes stride over the PU's 4x4 tiles.  x265's
 * satd of every shape is the sum over 4x4 tiles of the halved 4x4 Hadamard abs-sum (satd_4x4 for the widths 4 and 12,
 * satd_8x4 = two such tiles whose packed sum is halved: each tile's abs-sum is even, so halving per tile is the same). */
template <typename P>
__global__ void __launch_bounds__(256) pixelcmp_pu_kernel(int satd, int w, int h, const P* __restrict__ A, int64_t strideA, const P* __restrict__ B, int64_t strideB,
                                                           int n, const int64_t* __restrict__ offA, const int64_t* __restrict__ offB, int* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (idx >= n) return;
    const P* a = A + offA[idx];
    const P* b = B + offB[idx];
    const int tx = w >> 2, tiles = tx * (h >> 2);
    int acc = 0;
    for (int t = lane; t < tiles; t += 32)
    {
        const int x = (t % tx) * 4, y = (t / tx) * 4;
        typename Px<P>::Row4 fa[4], fb[4];
#pragma unroll
        for (int i = 0; i < 4; i++)
        {
            fa[i] = Px<P>::load(a + (int64_t)(y + i) * strideA + x);
            fb[i] = Px<P>::load(b + (int64_t)(y + i) * strideB + x);
        }
        acc += satd ? (satd4x4_abs<P>(fa, fb) >> 1) : sad4x4<P>(fa, fb);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[idx] = acc;
}

/* n block pairs at arbitrary sample offsets; one quad per pair (kind 3 = sa8d 16x16: four passes) */
template <typename P>
__global__ void __launch_bounds__(256) pixelcmp_batch_kernel(int kind, const P* __restrict__ A, int64_t strideA, const P* __restrict__ B, int64_t strideB,
                                                              int n, const int64_t* __restrict__ offA, const int64_t* __restrict__ offB, int* __restrict__ out)
{
    const int lane = threadIdx.x & 31;
    const int sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int idx = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const int valid = idx < n;
    const P* a = A + (valid ? offA[idx] : 0);
    const P* b = B + (valid ? offB[idx] : 0);
    int result = 0;
    const int nSub = kind == X265CU_SA8D_16x16_K ? 4 : 1;
    for (int s8 = 0; s8 < nSub; s8++)
    {
        const int ox = (s8 & 1) * 8, oy = (s8 >> 1) * 8;
        typename Px<P>::Row4 fa[4], fb[4];
#pragma unroll
        for (int y = 0; y < 4; y++)
        {
            if (valid)
            {
                fa[y] = Px<P>::load(a + (int64_t)(oy + by + y) * strideA + ox + bx);
                fb[y] = Px<P>::load(b + (int64_t)(oy + by + y) * strideB + ox + bx);
            }
            else { fa[y] = fb[y] = typename Px<P>::Row4(); }
        }
        if (kind == 0)
            result += quad_sum(sad4x4<P>(fa, fb));
        else if (kind == 1)
            result += quad_sum(satd4x4_abs<P>(fa, fb)) >> 1;
        else
        {
            int d[4][4];
#pragma unroll
            for (int y = 0; y < 4; y++)
            {
                int u[4], v[4];
                Px<P>::unpack(fa[y], u); Px<P>::unpack(fb[y], v);
#pragma unroll
                for (int x = 0; x < 4; x++) d[y][x] = u[x] - v[x];
            }
            result += quad_sum(sa8d_quad_abs(d, sub));
        }
    }
    if (kind >= 2) result = (result + 2) >> 2;
    if (valid && sub == 0) out[idx] = result;
}

/* 8 samples of a row in one load (8 or 16 bytes), as the two 4-sample halves the 4x4 measures take */
template <typename P> struct Wide8;
template <> struct Wide8<uint8_t>
{
    static __device__ __forceinline__ void load(const uint8_t* p, Px<uint8_t>::Row4& a, Px<uint8_t>::Row4& b)
    {
        const uint2 w = __ldg((const uint2*)p);
        a.v = w.x; b.v = w.y;
    }
};
template <> struct Wide8<uint16_t>
{
    static __device__ __forceinline__ void load(const uint16_t* p, Px<uint16_t>::Row4& a, Px<uint16_t>::Row4& b)
    {
        const uint4 w = __ldg((const uint4*)p);
        a.lo = w.x; a.hi = w.y; b.lo = w.z; b.hi = w.w;
    }
};

/* SAD / SATD of every aligned 8x8 block of plane 0 of pairs of frames, wide form: a lane takes the upper or lower
 * 8x4 half of a CU (two 4x4 blocks, one 8- or 16-byte load per row and plane), a warp 16 horizontally adjacent CUs
 * per group and GROUPS groups at once, all loads of an iteration in flight before the first is consumed
 * (8-bit, GROUPS = 2: 16 x 8 bytes per lane).  Half as many load instructions per sample as the quad form below and
 * twice the bytes in flight per thread; the quad form stays for SA8D, whose 8x8 transform needs the quad layout. */
template <typename P, int GROUPS>
__global__ void __launch_bounds__(256) pixelcmp_frames_wide_kernel(int kind, const void* const* __restrict__ planes, GeomDev g, int* __restrict__ out)
{
    typedef typename Px<P>::Row4 Row4;
    const P* A = (const P*)planes[2 * blockIdx.y];
    const P* B = (const P*)planes[2 * blockIdx.y + 1];
    out += (int64_t)blockIdx.y * g.nCU;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const int c = lane >> 1, half = lane & 1;
    for (int base = (blockIdx.x * nWarps + warp) * 16 * GROUPS; base < g.nCU; base += gridDim.x * nWarps * 16 * GROUPS)
    {
        Row4 fa[GROUPS][2][4], fb[GROUPS][2][4];
#pragma unroll
        for (int gi = 0; gi < GROUPS; gi++)
        {
            const int mb = base + gi * 16 + c;
            const int valid = mb < g.nCU;
            const int cuX = valid ? mb % g.wCU : 0, cuY = valid ? mb / g.wCU : 0;
            const int64_t off = (int64_t)(8 * cuY + 4 * half) * g.stride + 8 * cuX;
#pragma unroll
            for (int y = 0; y < 4; y++)
            {
                Wide8<P>::load(A + off + (int64_t)y * g.stride, fa[gi][0][y], fa[gi][1][y]);
                Wide8<P>::load(B + off + (int64_t)y * g.stride, fb[gi][0][y], fb[gi][1][y]);
            }
        }
#pragma unroll
        for (int gi = 0; gi < GROUPS; gi++)
        {
            const int mb = base + gi * 16 + c;
            int v = kind == 0 ? sad4x4<P>(fa[gi][0], fb[gi][0]) + sad4x4<P>(fa[gi][1], fb[gi][1])
                              : satd4x4_abs<P>(fa[gi][0], fb[gi][0]) + satd4x4_abs<P>(fa[gi][1], fb[gi][1]);
            v += __shfl_xor_sync(FULL_MASK, v, 1);
            if (kind == 1) v >>= 1;
            if (mb < g.nCU && half == 0) out[mb] = v;
        }
    }
}

/* every aligned 8x8 block of plane 0 of pairs of frames (the HBM-bound "SATD Gpix/s" kernel):
 * a warp covers 8 horizontally adjacent CUs, so each load instruction touches full 32-byte sectors.
 * grid = (blocks, pairs); planes[2 * pair], planes[2 * pair + 1] = sample (0,0) of the two planes. */
template <typename P>
__global__ void __launch_bounds__(256) pixelcmp_frames_kernel(int kind, const void* const* __restrict__ planes, GeomDev g, int* __restrict__ out)
{
    const P* A = (const P*)planes[2 * blockIdx.y];
    const P* B = (const P*)planes[2 * blockIdx.y + 1];
    out += (int64_t)blockIdx.y * g.nCU;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    for (int base = (blockIdx.x * nWarps + warp) * 8; base < g.nCU; base += gridDim.x * nWarps * 8)
    {
        const int mb = base + q;
        const int valid = mb < g.nCU;
        const int cuX = valid ? mb % g.wCU : 0, cuY = valid ? mb / g.wCU : 0;
        const int64_t off = (int64_t)(8 * cuY + by) * g.stride + 8 * cuX + bx;
        typename Px<P>::Row4 fa[4], fb[4];
#pragma unroll
        for (int y = 0; y < 4; y++)
        {
            fa[y] = Px<P>::load_aligned(A + off + (int64_t)y * g.stride);
            fb[y] = Px<P>::load_aligned(B + off + (int64_t)y * g.stride);
        }
        int result;
        if (kind == 0)
            result = quad_sum(sad4x4<P>(fa, fb));
        else if (kind == 1)
            result = quad_sum(satd4x4_abs<P>(fa, fb)) >> 1;
        else
        {
            int d[4][4];
#pragma unroll
            for (int y = 0; y < 4; y++)
            {
                int u[4], v[4];
                Px<P>::unpack(fa[y], u); Px<P>::unpack(fb[y], v);
#pragma unroll
                for (int x = 0; x < 4; x++) d[y][x] = u[x] - v[x];
            }
            result = (quad_sum(sa8d_quad_abs(d, sub)) + 2) >> 2;
        }
        if (valid && sub == 0) out[mb] = result;
    }
}

/* integer-pipe micro-benchmark (see x265cu_int_peak): 16 independent accumulator chains per thread */
__global__ void __launch_bounds__(256) int_peak_kernel(int mode, int iters, unsigned int* __restrict__ sink)
{
    unsigned int a[16];
#pragma unroll
    for (int k = 0; k < 16; k++) a[k] = threadIdx.x * 16 + k;
    unsigned int x = blockIdx.x * 0x9E3779B1u + threadIdx.x;
    if (mode == 0)
    {
        for (int i = 0; i < iters; i++)
        {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = __vsadu4(a[k], x) + a[k];
            x += 0x01010101u;
        }
    }
    else
    {
        for (int i = 0; i < iters; i++)
        {
#pragma unroll
            for (int k = 0; k < 16; k++) a[k] = a[k] + x + (unsigned int)k;
            x ^= a[0];
        }
    }
    unsigned int r = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) r ^= a[k];
    if (r == 0xDEADBEEFu) sink[0] = r;
}

/* ===========================================================================================
 * frame_var: pixel_var<16> (luma) + pixel_var<8> (Cb, Cr) per 16x16 block, acEnergyCu
 * (slicetype.cpp:48-93).  Wide forms (16-byte aligned luma rows, 8-byte aligned chroma rows).  PAIR
 * (the batch kernel's default): a pair of lanes per 16x16 block, a warp per 16 consecutive blocks; a
 * lane reads 8 luma rows of its block with one 16-byte load each (two at 16 bit) and 4 rows of Cb and
 * of Cr with one 8-byte load each, all issued before the first is consumed; every load instruction
 * of the warp covers whole 128-byte lines of the rows it touches (256 bytes of luma, 128 of chroma);
 * sums with the packed-byte instructions (__vsadu4 against 0, dp4a of a word with itself), a block's
 * six sums meet in one shuffle step.  Measured at 4K (48 frames, 7 launches): 0.199 ms against 0.287
 * ms for the QUAD form (a quad per block, a warp per 8 blocks, 4 + 2 rows per lane: 64-byte pieces
 * of the chroma rows per instruction), which the single-frame kernel keeps.
 * Otherwise: one warp per block, 64-bit or scalar loads.  Groups / blocks are strided over a grid
 * sized to the GPU; the six frame sums (wp_sum[0..2], wp_ssd[0..2]) are kept per lane, reduced per
 * CTA in shared memory and added to sums6 with six atomics per CTA.
 * =========================================================================================== */
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
template <typename P> struct VarAcc;
template <> struct VarAcc<uint8_t>
{
    static __device__ __forceinline__ void add(uint32_t w, unsigned int& sum, unsigned int& sqr) { sum += __vsadu4(w, 0); sqr = __dp4a(w, w, sqr); }
};
template <> struct VarAcc<uint16_t>
{
    static __device__ __forceinline__ void add(uint32_t w, unsigned int& sum, unsigned int& sqr)
    {
        const unsigned int lo = w & 0xffffu, hi = w >> 16;
        sum += lo + hi; sqr += lo * lo + hi * hi;
    }
};
__device__ __forceinline__ unsigned int quad_sum_u(unsigned int v)
{
    v += __shfl_xor_sync(FULL_MASK, v, 1);
    v += __shfl_xor_sync(FULL_MASK, v, 2);
    return v;
}
template <typename P, bool PAIR = false>
__device__ __forceinline__ void frame_var_body(const P* __restrict__ y, int64_t ys, const P* __restrict__ u, const P* __restrict__ v, int64_t cs,
                                               int blocksX, int blocksY, unsigned int* __restrict__ energy, unsigned long long* __restrict__ sums6,
                                               P* __restrict__ uKeep = NULL, P* __restrict__ vKeep = NULL)
{
    __shared__ unsigned long long sAcc[6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, warpsPerCta = blockDim.x >> 5;
    if (threadIdx.x < 6) sAcc[threadIdx.x] = 0;
    __syncthreads();
    unsigned long long acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0, acc4 = 0, acc5 = 0;
    const bool packed = sizeof(P) == 1 && (((uintptr_t)y | (uintptr_t)(ys * (int64_t)sizeof(P))) & 7) == 0;
    const bool wide = (((uintptr_t)y | (uintptr_t)(ys * (int64_t)sizeof(P))) & 15) == 0 &&
                      (!(u && v) || ((((uintptr_t)u | (uintptr_t)v | (uintptr_t)(cs * (int64_t)sizeof(P))) & (8 * sizeof(P) - 1)) == 0 &&
                                     (!uKeep || (((uintptr_t)uKeep | (uintptr_t)vKeep) & (8 * sizeof(P) - 1)) == 0)));
    const int nBlk = blocksX * blocksY;
    if (PAIR && wide)
    {
        /* a PAIR of lanes per 16x16 block, a warp per 16 consecutive blocks: a lane reads 8 luma rows (16-byte loads) and 4 rows
         * of Cb and of Cr; every load instruction of the warp covers whole 128-byte lines of each row it touches */
        const int q = lane >> 1, sub = lane & 1;
        for (int grp = blockIdx.x * warpsPerCta + warp; grp * 16 < nBlk; grp += gridDim.x * warpsPerCta)
        {
            const int blk = grp * 16 + q;
            const bool valid = blk < nBlk;
            const int bxi = valid ? blk % blocksX : 0, byi = valid ? blk / blocksX : 0;
            unsigned int sum = 0, sqr = 0, s1 = 0, q1 = 0, s2 = 0, q2 = 0;
            if (valid)
            {
                const P* p = y + (int64_t)(16 * byi + 8 * sub) * ys + 16 * bxi;
                uint4 w[8 * (int)sizeof(P)];
#pragma unroll
                for (int r = 0; r < 8; r++)
#pragma unroll
                    for (int h = 0; h < (int)sizeof(P); h++)
                        w[r * (int)sizeof(P) + h] = __ldg((const uint4*)(p + (int64_t)r * ys) + h);
                if (u && v)
                {
                    const int64_t co = (int64_t)(8 * byi + 4 * sub) * cs + 8 * bxi;
                    const int64_t ko = (int64_t)(8 * byi + 4 * sub) * (8 * blocksX) + 8 * bxi;
                    if (sizeof(P) == 1)
                    {
                        uint2 a[4], b[4];
#pragma unroll
                        for (int r = 0; r < 4; r++) { a[r] = __ldg((const uint2*)(u + co + r * cs)); b[r] = __ldg((const uint2*)(v + co + r * cs)); }
#pragma unroll
                        for (int r = 0; r < 4; r++)
                        {
                            VarAcc<P>::add(a[r].x, s1, q1); VarAcc<P>::add(a[r].y, s1, q1);
                            VarAcc<P>::add(b[r].x, s2, q2); VarAcc<P>::add(b[r].y, s2, q2);
                            if (uKeep) { *(uint2*)(uKeep + ko + r * (8 * blocksX)) = a[r]; *(uint2*)(vKeep + ko + r * (8 * blocksX)) = b[r]; }
                        }
                    }
                    else
                    {
#pragma unroll
                        for (int r = 0; r < 4; r++)
                        {
                            const uint4 a = __ldg((const uint4*)(u + co + r * cs)), b = __ldg((const uint4*)(v + co + r * cs));
                            VarAcc<P>::add(a.x, s1, q1); VarAcc<P>::add(a.y, s1, q1); VarAcc<P>::add(a.z, s1, q1); VarAcc<P>::add(a.w, s1, q1);
                            VarAcc<P>::add(b.x, s2, q2); VarAcc<P>::add(b.y, s2, q2); VarAcc<P>::add(b.z, s2, q2); VarAcc<P>::add(b.w, s2, q2);
                            if (uKeep) { *(uint4*)(uKeep + ko + r * (8 * blocksX)) = a; *(uint4*)(vKeep + ko + r * (8 * blocksX)) = b; }
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < 8 * (int)sizeof(P); k++)
                {
                    VarAcc<P>::add(w[k].x, sum, sqr); VarAcc<P>::add(w[k].y, sum, sqr); VarAcc<P>::add(w[k].z, sum, sqr); VarAcc<P>::add(w[k].w, sum, sqr);
                }
            }
            acc0 += sum; acc3 += sqr; acc1 += s1; acc4 += q1; acc2 += s2; acc5 += q2;
            sum += __shfl_xor_sync(FULL_MASK, sum, 1); sqr += __shfl_xor_sync(FULL_MASK, sqr, 1);
            unsigned int var = sqr - (unsigned int)(((unsigned long long)sum * sum) >> 8);
            if (u && v)
            {
                s1 += __shfl_xor_sync(FULL_MASK, s1, 1); q1 += __shfl_xor_sync(FULL_MASK, q1, 1);
                s2 += __shfl_xor_sync(FULL_MASK, s2, 1); q2 += __shfl_xor_sync(FULL_MASK, q2, 1);
                var += q1 - (unsigned int)(((unsigned long long)s1 * s1) >> 6);
                var += q2 - (unsigned int)(((unsigned long long)s2 * s2) >> 6);
            }
            if (valid && sub == 0) energy[blk] = var;
        }
        acc0 = (unsigned long long)warp_sum_u64(acc0); acc3 = warp_sum_u64(acc3);
        if (u && v) { acc1 = warp_sum_u64(acc1); acc4 = warp_sum_u64(acc4); acc2 = warp_sum_u64(acc2); acc5 = warp_sum_u64(acc5); }
    }
    else if (!PAIR && wide)
    {
        const int q = lane >> 2, sub = lane & 3;
        for (int grp = blockIdx.x * warpsPerCta + warp; grp * 8 < nBlk; grp += gridDim.x * warpsPerCta)
        {
            const int blk = grp * 8 + q;
            const bool valid = blk < nBlk;
            const int bxi = valid ? blk % blocksX : 0, byi = valid ? blk / blocksX : 0;
            unsigned int sum = 0, sqr = 0, s1 = 0, q1 = 0, s2 = 0, q2 = 0;
            if (valid)
            {
                const P* p = y + (int64_t)(16 * byi + 4 * sub) * ys + 16 * bxi;
#pragma unroll
                for (int r = 0; r < 4; r++)
#pragma unroll
                    for (int h = 0; h < (int)sizeof(P); h++)
                    {
                        const uint4 w = __ldg((const uint4*)(p + (int64_t)r * ys) + h);
                        VarAcc<P>::add(w.x, sum, sqr); VarAcc<P>::add(w.y, sum, sqr); VarAcc<P>::add(w.z, sum, sqr); VarAcc<P>::add(w.w, sum, sqr);
                    }
                if (u && v)
                {
#pragma unroll
                    for (int r = 0; r < 2; r++)
                    {
                        const int64_t co = (int64_t)(8 * byi + 2 * sub + r) * cs + 8 * bxi;
                        const int64_t ko = (int64_t)(8 * byi + 2 * sub + r) * (8 * blocksX) + 8 * bxi;
                        if (sizeof(P) == 1)
                        {
                            const uint2 a = __ldg((const uint2*)(u + co)), b = __ldg((const uint2*)(v + co));
                            VarAcc<P>::add(a.x, s1, q1); VarAcc<P>::add(a.y, s1, q1);
                            VarAcc<P>::add(b.x, s2, q2); VarAcc<P>::add(b.y, s2, q2);
                            if (uKeep) { *(uint2*)(uKeep + ko) = a; *(uint2*)(vKeep + ko) = b; }
                        }
                        else
                        {
                            const uint4 a = __ldg((const uint4*)(u + co)), b = __ldg((const uint4*)(v + co));
                            VarAcc<P>::add(a.x, s1, q1); VarAcc<P>::add(a.y, s1, q1); VarAcc<P>::add(a.z, s1, q1); VarAcc<P>::add(a.w, s1, q1);
                            VarAcc<P>::add(b.x, s2, q2); VarAcc<P>::add(b.y, s2, q2); VarAcc<P>::add(b.z, s2, q2); VarAcc<P>::add(b.w, s2, q2);
                            if (uKeep) { *(uint4*)(uKeep + ko) = a; *(uint4*)(vKeep + ko) = b; }
                        }
                    }
                }
            }
            acc0 += sum; acc3 += sqr; acc1 += s1; acc4 += q1; acc2 += s2; acc5 += q2;
            sum = quad_sum_u(sum); sqr = quad_sum_u(sqr);
            unsigned int var = sqr - (unsigned int)(((unsigned long long)sum * sum) >> 8);
            if (u && v)
            {
                s1 = quad_sum_u(s1); q1 = quad_sum_u(q1); s2 = quad_sum_u(s2); q2 = quad_sum_u(q2);
                var += q1 - (unsigned int)(((unsigned long long)s1 * s1) >> 6);
                var += q2 - (unsigned int)(((unsigned long long)s2 * s2) >> 6);
            }
            if (valid && sub == 0) energy[blk] = var;
        }
        acc0 = (unsigned long long)warp_sum_u64(acc0); acc3 = warp_sum_u64(acc3);
        if (u && v) { acc1 = warp_sum_u64(acc1); acc4 = warp_sum_u64(acc4); acc2 = warp_sum_u64(acc2); acc5 = warp_sum_u64(acc5); }
    }
    else
    for (int blk = blockIdx.x * warpsPerCta + warp; blk < nBlk; blk += gridDim.x * warpsPerCta)
    {
        const int bxi = blk % blocksX, byi = blk / blocksX;
        /* luma: 256 samples, 8 per lane */
        unsigned int sum = 0, sqr = 0;
        {
            const P* p = y + (int64_t)(16 * byi + (lane >> 1)) * ys + 16 * bxi + (lane & 1) * 8;
            if (packed)
            {
                const uint2 w = __ldg((const uint2*)p);
                sum = __vsadu4(w.x, 0) + __vsadu4(w.y, 0);
                sqr = __dp4a(w.x, w.x, __dp4a(w.y, w.y, 0u));
            }
            else
            {
#pragma unroll
                for (int i = 0; i < 8; i++) { unsigned int t = p[i]; sum += t; sqr += t * t; }
            }
        }
        sum = (unsigned int)warp_sum((int)sum); sqr = (unsigned int)warp_sum((int)sqr);
        unsigned int var = sqr - (unsigned int)(((unsigned long long)sum * sum) >> 8);
        unsigned int s1 = 0, q1 = 0, s2 = 0, q2 = 0;
        if (u && v)
        {
            const int64_t co = (int64_t)(8 * byi + (lane >> 2)) * cs + 8 * bxi + (lane & 3) * 2;
#pragma unroll
            for (int i = 0; i < 2; i++)
            {
                unsigned int a = u[co + i], b = v[co + i];
                s1 += a; q1 += a * a; s2 += b; q2 += b * b;
                if (uKeep)
                {
                    /* every chroma sample passes through here once: leave a compact copy (pitch = 8 * blocksX) behind */
                    const int64_t ko = (int64_t)(8 * byi + (lane >> 2)) * (8 * blocksX) + 8 * bxi + (lane & 3) * 2 + i;
                    uKeep[ko] = (P)a; vKeep[ko] = (P)b;
                }
            }
            s1 = (unsigned int)warp_sum((int)s1); q1 = (unsigned int)warp_sum((int)q1);
            s2 = (unsigned int)warp_sum((int)s2); q2 = (unsigned int)warp_sum((int)q2);
            var += q1 - (unsigned int)(((unsigned long long)s1 * s1) >> 6);
            var += q2 - (unsigned int)(((unsigned long long)s2 * s2) >> 6);
        }
        if (lane == 0) energy[blk] = var;
        acc0 += sum; acc3 += sqr; acc1 += s1; acc4 += q1; acc2 += s2; acc5 += q2;
    }
    if (lane == 0)
    {
        atomicAdd(&sAcc[0], acc0); atomicAdd(&sAcc[3], acc3);
        if (u && v) { atomicAdd(&sAcc[1], acc1); atomicAdd(&sAcc[4], acc4); atomicAdd(&sAcc[2], acc2); atomicAdd(&sAcc[5], acc5); }
    }
    __syncthreads();
    if (threadIdx.x < 6 && sAcc[threadIdx.x]) atomicAdd(&sums6[threadIdx.x], sAcc[threadIdx.x]);
}

template <typename P>
__global__ void __launch_bounds__(256) frame_var_kernel(const P* __restrict__ y, int64_t ys, const P* __restrict__ u, const P* __restrict__ v, int64_t cs,
                                                         int blocksX, int blocksY, unsigned int* __restrict__ energy, unsigned long long* __restrict__ sums6,
                                                         P* __restrict__ uKeep, P* __restrict__ vKeep)
{
    frame_var_body<P>(y, ys, u, v, cs, blocksX, blocksY, energy, sums6, uKeep, vKeep);
}

template <typename P, bool PAIR>
__global__ void __launch_bounds__(256) frame_var_batch_kernel(VarBatch b, int blocksX, int blocksY)
{
    const int f = blockIdx.y;
    frame_var_body<P, PAIR>((const P*)b.y[f], b.ys[f], (const P*)b.u[f], (const P*)b.v[f], b.cs[f], blocksX, blocksY, b.energy[f], b.sums[f], (P*)b.uKeep[f], (P*)b.vKeep[f]);
}

#endif /* X265CU_KERNELS_CUH */
