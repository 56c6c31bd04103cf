/* x265cu_intra.cuh -- LookaheadTLD::lowresIntraEstimate (encoder/slicetype.cpp:230-336) on the GPU.
 *
 * One warp per 8x8 CU (CUs have no dependency on each other: all neighbours come from the padded
 * SOURCE plane).  The 33 neighbour samples and their [1 2 1] filtered copy (intraFilter<8>,
 * common/intrapred.cpp:31-51) live in shared memory.  As in the search kernel a quad measures one
 * candidate (here: one prediction mode) and each lane its 4x4 sub-block:
 *   pass 1: DC, planar, angular 5,10,...,30   (8 modes = 8 quads)
 *   pass 2: the six modes the two refinement rounds (best-2 / best+2, then -1 / +1 around the winner of that,
 *           slicetype.cpp:292-310) can ask for: best-3 .. best+3 without best; the two rounds are then decided on the
 *           measured costs in the reference's order (strict <, minus before plus)
 * Angular prediction (intra_pred_ang_c<8>, intrapred.cpp:102-204) is done in its "vertical" form
 * for every mode: horizontal modes swap the roles of the top and left neighbours and produce the
 * transposed block; since a 4x4 Hadamard abs-sum is invariant under transposition, the lane just
 * compares against its transposed source sub-block and no transposition is ever materialised.
 * Each quad first builds the mode's reference line (projected left samples, top-left, top,
 * top-right) in shared memory, so the per-sample work is two shared loads and one interpolation.
 */
#ifndef X265CU_INTRA_CUH
#define X265CU_INTRA_CUH

struct IntraOutDev
{
    int32_t* intraCost;
    uint8_t* intraMode;
    uint16_t* lowresCosts;    /* [0][0] */
    int32_t* rowSatds;        /* [0][0], zeroed before launch */
    unsigned long long* sums; /* costEst, costEstAq; zeroed before launch */
    const int32_t* invQ;      /* or NULL */
};

#define INTRA_EXT 32          /* reference line entries per quad: k = -9 .. 22 -> index k + 9 */

/* neighbour index after the horizontal-mode swap of top and left (intrapred.cpp:111-120) */
__device__ __forceinline__ int intra_swap(int hor, int k)
{
    return (hor && k > 0) ? (k <= 16 ? k + 16 : k - 16) : k;
}

#define INTRA_BATCH 8
struct IntraBatch { const void* plane0[INTRA_BATCH]; IntraOutDev o[INTRA_BATCH]; };

template <typename P>
__device__ __forceinline__ void intra_body(const P* __restrict__ plane0, const GeomDev& g, int lambda, int pixelMax, const IntraOutDev& o)
{
    __shared__ P sNb[8][2][36];
    __shared__ short sExt[8][8][INTRA_EXT];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cuXY = blockIdx.x * 8 + warp;
    if (cuXY >= g.nCU) return;
    const int cuX = cuXY % g.wCU, cuY = cuXY / g.wCU;
    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const P* pix = plane0 + (int64_t)(8 * cuY) * g.stride + 8 * cuX;
    P* s = sNb[warp][0];
    P* f = sNb[warp][1];
    short* ext = sExt[warp][q];
    /* neighbours: 17 samples of the row above from the top-left, 16 of the left column (slicetype.cpp:264-267) */
    {
        const P* tl = pix - g.stride - 1;
        if (lane < 17) s[lane] = tl[lane];
        if (lane < 16) s[17 + lane] = tl[(int64_t)(lane + 1) * g.stride];
        __syncwarp();
        for (int i = lane; i < 33; i += 32)
        {
            int v;
            if (i == 0) v = (2 * s[0] + s[1] + s[17] + 2) >> 2;
            else if (i == 16 || i == 32) v = s[i];
            else if (i == 17) v = (2 * s[17] + s[0] + s[18] + 2) >> 2;
            else v = (2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2;
            f[i] = (P)v;
        }
        __syncwarp();
    }
    int dcVal = 8;
#pragma unroll
    for (int i = 0; i < 8; i++) dcVal += s[1 + i] + s[17 + i];
    dcVal >>= 4;   /* dcVal / 16, dcVal >= 0 */

    /* source sub-block, plain (fe) and transposed position (ft): ft[r][c] = source(x = by + r... see below) */
    int fe[4][4];   /* fe[y][x] = source at (bx + x, by + y) */
    int ft[4][4];   /* ft[r][c] = source at (x = by + r, y = bx + c): the sub-block a horizontal mode's
                       vertical-form prediction rows r = bx'.. / cols c = by'.. must be compared with */
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        typename Px<P>::Row4 r = Px<P>::load_aligned(pix + (int64_t)(by + y) * g.stride + bx);
        Px<P>::unpack(r, fe[y]);
    }
    {
        /* transposed view: lane (bx, by) in a horizontal mode computes blk rows bx.., cols by.. and
         * needs source(x = bx + r, y = by + c) = fe[c][r] */
#pragma unroll
        for (int r = 0; r < 4; r++)
#pragma unroll
            for (int c = 0; c < 4; c++) ft[r][c] = fe[c][r];
    }

    int icost = LA_COST_MAX, ilow = 0, acost = LA_COST_MAX, alow = 4;
#pragma unroll 1
    for (int pass = 0; pass < 2; pass++)
    {
        int mode, valid = 1;
        if (pass == 0)
            mode = q == 0 ? 1 : (q == 1 ? 0 : 5 * (q - 1));          /* DC, planar, 5,10,...,30 */
        else
        {
            /* quads 0..5: alow -2, +2, -3, -1, +1, +3 (alow is one of 5,...,30: all within 2..33) */
            const int off = q == 0 ? -2 : q == 1 ? 2 : q == 2 ? -3 : q == 3 ? -1 : q == 4 ? 1 : 3;
            mode = alow + off;
            valid = q < 6;
            if (!valid) mode = 10;
        }
        int d[4][4];
        /* every quad builds a reference line (DC / planar quads for a dummy mode) so that the
         * warp-level barriers below are reached by all 32 lanes */
        const int amode = mode >= 2 ? mode : 10;
        /* intra_pred_ang_c<8>, intrapred.cpp:102-204; filtered neighbours only for modes 2, 18, 34 */
        const P* n = (amode == 2 || amode == 18 || amode == 34) ? f : s;
        const int hor = amode < 18;
        const int angleOffset = hor ? 10 - amode : amode - 26;
        const int mag = angleOffset < 0 ? -angleOffset : angleOffset;
        /* angleTable[8 + angleOffset] = sign * {0,2,5,9,13,17,21,26,32}[mag] */
        const int absAng = (int)((0x201A15110D090502ull >> (8 * ((mag - 1) & 7))) & 0xff);
        const int angle = mag == 0 ? 0 : (angleOffset < 0 ? -absAng : absAng);
        /* invAngleTable = {4096,1638,910,630,482,390,315,256}[mag - 1] (used only when angle < 0) */
        const int invAngle = mag == 1 ? 4096 : mag == 2 ? 1638 : mag == 3 ? 910 : mag == 4 ? 630 : mag == 5 ? 482 : mag == 6 ? 390 : mag == 7 ? 315 : 256;
        /* reference line of this mode: ref(k), k >= -1 -> neighbour k + 1 (top-left, top, top-right);
         * k < -1 -> projected left neighbour 16 + ((128 + (-1 - k) * invAngle) >> 8) */
#pragma unroll
        for (int j = 0; j < 8; j++)
        {
            const int k = sub * 8 + j - 9;
            int idx = k >= -1 ? k + 1 : 16 + ((128 + (-1 - k) * invAngle) >> 8);
            idx = idx > 32 ? 32 : idx;
            ext[sub * 8 + j] = (short)n[intra_swap(hor, idx)];
        }
        __syncwarp();
        if (mode == 1)
        {
            /* intra_pred_dc_c<8> + dcPredFilter, intrapred.cpp:53-85 (unfiltered neighbours) */
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++)
                {
                    const int X = bx + x, Y = by + y;
                    int v = dcVal;
                    if (X == 0 && Y == 0) v = (s[1] + s[17] + 2 * dcVal + 2) >> 2;
                    else if (Y == 0) v = (s[1 + X] + 3 * dcVal + 2) >> 2;
                    else if (X == 0) v = (s[17 + Y] + 3 * dcVal + 2) >> 2;
                    d[y][x] = fe[y][x] - v;
                }
        }
        else if (mode == 0)
        {
            /* planar_pred_c<3> on the filtered neighbours, intrapred.cpp:87-100 */
            const int topRight = f[9], bottomLeft = f[25];
#pragma unroll
            for (int y = 0; y < 4; y++)
#pragma unroll
                for (int x = 0; x < 4; x++)
                {
                    const int X = bx + x, Y = by + y;
                    d[y][x] = fe[y][x] - (((7 - X) * f[17 + Y] + (7 - Y) * f[1 + X] + (X + 1) * topRight + (Y + 1) * bottomLeft + 8) >> 4);
                }
        }
        else
        {
            /* vertical form: this lane computes blk rows r0.., cols c0.. */
            const int r0 = hor ? bx : by, c0 = hor ? by : bx;
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                const int pos = (r0 + r + 1) * angle;
                const int off = pos >> 5, frac = pos & 31;
                const short* e = ext + off + c0 + 9;
                int v0 = e[0], v1 = e[1], v2 = e[2], v3 = e[3], v4 = e[4];
                int p0 = ((32 - frac) * v0 + frac * v1 + 16) >> 5;
                int p1 = ((32 - frac) * v1 + frac * v2 + 16) >> 5;
                int p2 = ((32 - frac) * v2 + frac * v3 + 16) >> 5;
                int p3 = ((32 - frac) * v3 + frac * v4 + 16) >> 5;
                if (angle == 0 && c0 == 0)
                {
                    /* pure vertical / horizontal with the edge filter (bFilter = 1): first column */
                    int t = (int)n[intra_swap(hor, 1)] + (((int)n[intra_swap(hor, 17 + r0 + r)] - (int)n[0]) >> 1);
                    p0 = t < 0 ? 0 : (t > pixelMax ? pixelMax : t);
                }
                if (hor) { d[r][0] = ft[r][0] - p0; d[r][1] = ft[r][1] - p1; d[r][2] = ft[r][2] - p2; d[r][3] = ft[r][3] - p3; }
                else { d[r][0] = fe[r][0] - p0; d[r][1] = fe[r][1] - p1; d[r][2] = fe[r][2] - p2; d[r][3] = fe[r][3] - p3; }
            }
        }
        __syncwarp();   /* the reference lines are rebuilt in the next pass */
        int cost = quad_sum(hadamard4x4_abs(d)) >> 1;
        if (!valid) cost = LA_COST_MAX;
        int c[8];
#pragma unroll
        for (int k = 0; k < 8; k++) c[k] = __shfl_sync(FULL_MASK, cost, 4 * k);
        if (pass == 0)
        {
            if (c[0] < icost) { icost = c[0]; ilow = 1; }
            if (c[1] < icost) { icost = c[1]; ilow = 0; }
#pragma unroll
            for (int k = 2; k < 8; k++)
                if (c[k] < acost) { acost = c[k]; alow = 5 * (k - 1); }
        }
        else
        {
            const int a = alow;
            if (c[0] < acost) { acost = c[0]; alow = a - 2; }
            if (c[1] < acost) { acost = c[1]; alow = a + 2; }
            const int a2 = alow;
            const int cm = a2 < a ? c[2] : (a2 == a ? c[3] : c[4]);     /* cost of a2 - 1 */
            const int cp = a2 < a ? c[3] : (a2 == a ? c[4] : c[5]);     /* cost of a2 + 1 */
            if (cm < acost) { acost = cm; alow = a2 - 1; }
            if (cp < acost) { acost = cp; alow = a2 + 1; }
        }
    }
    if (acost < icost) { icost = acost; ilow = alow; }
    icost += 5 * lambda + 4;
    if (lane == 0)
    {
        int capped = icost < LA_LOWRES_COST_MASK ? icost : LA_LOWRES_COST_MASK;
        o.lowresCosts[cuXY] = (uint16_t)capped;
        o.intraCost[cuXY] = icost;
        o.intraMode[cuXY] = (uint8_t)ilow;
        int scored = (cuX > 0 && cuX < g.wCU - 1 && cuY > 0 && cuY < g.hCU - 1) || g.wCU <= 2 || g.hCU <= 2;
        int icostAq = (scored && o.invQ) ? ((icost * o.invQ[cuXY] + 128) >> 8) : icost;
        if (scored)
        {
            atomicAdd(&o.sums[0], (unsigned long long)icost);
            atomicAdd(&o.sums[1], (unsigned long long)icostAq);
        }
        atomicAdd(&o.rowSatds[cuY], icostAq);
    }
}

template <typename P>
__global__ void __launch_bounds__(256) intra_kernel(const P* __restrict__ plane0, GeomDev g, int lambda, int pixelMax, IntraOutDev o)
{
    intra_body<P>(plane0, g, lambda, pixelMax, o);
}

/* the frames of a pre-lookahead list in one launch: blockIdx.y = frame */
template <typename P>
__global__ void __launch_bounds__(256) intra_batch_kernel(IntraBatch b, GeomDev g, int lambda, int pixelMax)
{
    intra_body<P>((const P*)b.plane0[blockIdx.y], g, lambda, pixelMax, b.o[blockIdx.y]);
}

#endif /* X265CU_INTRA_CUH */
