/* x265cu_wp.cuh -- pixel work of the explicit weighted-prediction analysis on the GPU (sm_100a).
 *
 * encoder/weightPrediction.cpp decides, per inter slice, list and plane, whether a (scale, offset) weight pays: it builds a
 * motion-compensated copy of the reference with the lookahead's lowres vectors (mcLuma :59-90 on the lowres planes, mcChroma
 * :92-166 on the full-resolution chroma planes with the 4-tap interpolation filters of common/ipfilter.cpp) and measures
 * sum(min(SATD 8x8, intraCost)) of the weighted copy for a sweep of candidate weights (weightCost :168-220, weight_pp_c
 * common/pixel.cpp:463-488).  The float guesses, the sweep order, its early exits and the 0.998 acceptance test stay the
 * host's (x265's own code); the three pixel loops run here on planes that are already resident: the lowres planes of the
 * frame slots, and a compact copy of the source chroma planes that frame_var_kernel leaves behind (it reads every chroma
 * sample anyway).  Nothing is copied back but one 32-bit cost per candidate, and the host no longer needs the lowres planes.
 */
#ifndef X265CU_WP_CUH
#define X265CU_WP_CUH

/* g_chromaFilter (common/constants.cpp:247-257) */
__constant__ short c_wpChromaFilter[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };
__device__ __forceinline__ int wp_chroma_tap(int frac, int k) { return c_wpChromaFilter[frac][k]; }

/* mcLuma (weightPrediction.cpp:59-90): out = lowres reference motion-compensated CU by CU with the lookahead's vectors,
 * each clipped to the picture + 8 samples; out has the lowres stride and no borders.  A quad per CU, a lane per 4x4. */
template <typename P>
__global__ void __launch_bounds__(256) wp_mc_luma_kernel(const P* __restrict__ refPlane0, GeomDev g, const int* __restrict__ mvs, P* __restrict__ out)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const int cu = (blockIdx.x * (blockDim.x >> 5) + warp) * 8 + q;
    if (cu >= g.nCU) return;
    const int cuX = cu % g.wCU, cuY = cu / g.wCU;
    const int x = 8 * cuX, y = 8 * cuY;
    const int m = mvs[cu];
    int mvx = la_mv_x(m), mvy = la_mv_y(m);
    const int minx = (int)(int16_t)((-x - 8) * 4), maxx = (int)(int16_t)((g.width - x - 1 + 8) * 4);
    const int miny = (int)(int16_t)((-y - 8) * 4), maxy = (int)(int16_t)((g.lines - y - 1 + 8) * 4);
    mvx = mvx < minx ? minx : (mvx > maxx ? maxx : mvx);
    mvy = mvy < miny ? miny : (mvy > maxy ? maxy : mvy);
    RefPlanes<P> r = { refPlane0, g.planeSize, g.stride };
    typename Px<P>::Row4 rows[4];
    mc_fetch4x4<P>(r, x + bx, y + by, mvx, mvy, rows);
#pragma unroll
    for (int i = 0; i < 4; i++)
    {
        int v[4];
        Px<P>::unpack(rows[i], v);
        P* d = out + (int64_t)(y + by + i) * g.stride + x + bx;
        d[0] = (P)v[0]; d[1] = (P)v[1]; d[2] = (P)v[2]; d[3] = (P)v[3];
    }
}

/* mcChroma (weightPrediction.cpp:92-166), 4:2:0: a thread per output sample of the (width x height) analysis area.
 * src = the reference's compact chroma plane (cW x cH valid samples, pitch cP); reads beyond it replicate the edge, which is
 * what extendPicBorder (:296-304) provides.  The block loop's availability test compares SAMPLE positions with the lowres
 * CU counts and indexes the vectors with y * widthInCU + x / 8 (:113-121): restated literally. */
template <typename P>
__global__ void __launch_bounds__(256) wp_mc_chroma_kernel(const P* __restrict__ src, int cP, int cW, int cH, const int* __restrict__ mvs, int wCU, int hCU,
                                                           int width, int height, P* __restrict__ out, int outPitch, int depth)
{
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (x >= width || y >= height) return;
    const int X0 = x & ~7, Y0 = y & ~7;
#define WP_S(xx, yy) ((int)src[(int64_t)((yy) < 0 ? 0 : ((yy) > cH - 1 ? cH - 1 : (yy))) * cP + ((xx) < 0 ? 0 : ((xx) > cW - 1 ? cW - 1 : (xx)))])
    int val;
    if (X0 < wCU && Y0 < hCU)
    {
        const int m = mvs[Y0 * wCU + (X0 >> 3)];
        /* lowres MV -> full-resolution MV -> chroma MV: mv <<= 1; mv.x >>= hshift; mv.y >>= vshift (int16 fields) */
        int mvx = (int)(int16_t)(la_mv_x(m) << 1) >> 1, mvy = (int)(int16_t)(la_mv_y(m) << 1) >> 1;
        const int minx = (int)(int16_t)((-X0 - 8) * 4), maxx = (int)(int16_t)((width - X0 - 1 + 8) * 4);
        const int miny = (int)(int16_t)((-Y0 - 8) * 4), maxy = (int)(int16_t)((height - Y0 - 1 + 8) * 4);
        mvx = mvx < minx ? minx : (mvx > maxx ? maxx : mvx);
        mvy = mvy < miny ? miny : (mvy > maxy ? maxy : mvy);
        const int sx = x + (mvx >> 2), sy = y + (mvy >> 2);
        const int xFrac = mvx & 7, yFrac = mvy & 7;
        const int maxVal = (1 << depth) - 1;
        if (!(xFrac | yFrac))
            val = WP_S(sx, sy);
        else if (!yFrac)
        {
            /* interp_horiz_pp_c<4> (ipfilter.cpp:80-119) */
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) sum += WP_S(sx - 1 + k, sy) * wp_chroma_tap(xFrac, k);
            val = (int)(int16_t)((sum + 32) >> 6);
            val = val < 0 ? 0 : (val > maxVal ? maxVal : val);
        }
        else if (!xFrac)
        {
            /* interp_vert_pp_c<4> (ipfilter.cpp:166-205) */
            int sum = 0;
#pragma unroll
            for (int k = 0; k < 4; k++) sum += WP_S(sx, sy - 1 + k) * wp_chroma_tap(yFrac, k);
            val = (int)(int16_t)((sum + 32) >> 6);
            val = val < 0 ? 0 : (val > maxVal ? maxVal : val);
        }
        else
        {
            /* interp_horiz_ps_c<4> with row extension (:121-163), then interp_vert_sp_c<4> (:245-284) */
            const int headRoom = 14 - depth;
            const int shiftH = 6 - headRoom, offH = -(8192 << shiftH);
            const int shiftV = 6 + headRoom, offV = (1 << (shiftV - 1)) + (8192 << 6);
            int sum = 0;
#pragma unroll
            for (int r = 0; r < 4; r++)
            {
                int h = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) h += WP_S(sx - 1 + k, sy - 1 + r) * wp_chroma_tap(xFrac, k);
                const int imm = (int)(int16_t)((h + offH) >> shiftH);
                sum += imm * wp_chroma_tap(yFrac, r);
            }
            val = (int)(int16_t)((sum + offV) >> shiftV);
            val = val < 0 ? 0 : (val > maxVal ? maxVal : val);
        }
    }
    else
        val = WP_S(x, y);
#undef WP_S
    out[(int64_t)y * outPitch + x] = (P)val;
}

/* weightCost (weightPrediction.cpp:168-220): sum over the 8x8 blocks of a plane of SATD(weighted reference, source),
 * limited by intraCost[block] for luma.  The weighted copy (weight_pp_c) is never materialised.  grid = (blocks of warps,
 * candidates): one launch measures a whole list of candidate weights. */
struct WpCand { int weighted, scale, round, shift, offset; };
struct WpCostArgs
{
    const void* fenc; const void* ref;
    int fencStride, refStride;
    int wBlk, nBlk;              /* 8x8 blocks per row, blocks in all */
    const int* intraCost;        /* or NULL (chroma) */
};

template <typename P>
__global__ void __launch_bounds__(256) wp_cost_kernel(WpCostArgs a, const WpCand* __restrict__ cands, unsigned int* __restrict__ costs, int correction, int pixelMax)
{
    const WpCand w = cands[blockIdx.y];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const int q = lane >> 2, sub = lane & 3, bx = (sub & 1) * 4, by = (sub >> 1) * 4;
    const P* fenc = (const P*)a.fenc;
    const P* ref = (const P*)a.ref;
    unsigned int acc = 0;
    for (int base = (blockIdx.x * nWarps + warp) * 8; base < a.nBlk; base += gridDim.x * nWarps * 8)
    {
        const int mb = base + q;
        const int valid = mb < a.nBlk;
        int part = 0;
        if (valid)
        {
            const int bX = mb % a.wBlk, bY = mb / a.wBlk;
            const int64_t offR = (int64_t)(8 * bY + by) * a.refStride + 8 * bX + bx;
            const int64_t offF = (int64_t)(8 * bY + by) * a.fencStride + 8 * bX + bx;
            int d[4][4];
#pragma unroll
            for (int y = 0; y < 4; y++)
            {
                int r4[4], f4[4];
                Px<P>::unpack(Px<P>::load(ref + offR + (int64_t)y * a.refStride), r4);
                Px<P>::unpack(Px<P>::load(fenc + offF + (int64_t)y * a.fencStride), f4);
#pragma unroll
                for (int x = 0; x < 4; x++)
                {
                    int v = r4[x];
                    if (w.weighted)
                    {
                        const int val = (int)(int16_t)(v << correction);
                        v = ((w.scale * val + w.round) >> w.shift) + w.offset;
                        v = v < 0 ? 0 : (v > pixelMax ? pixelMax : v);
                    }
                    d[y][x] = v - f4[x];
                }
            }
            part = hadamard4x4_abs(d);
        }
        const int satd = quad_sum(part) >> 1;
        if (valid && sub == 0)
        {
            if (a.intraCost)
            {
                const int ic = a.intraCost[mb];
                acc += (unsigned int)(satd < ic ? satd : ic);
            }
            else
                acc += (unsigned int)satd;
        }
    }
    acc = (unsigned int)warp_sum((int)acc);
    if (lane == 0 && acc)
        atomicAdd(&costs[blockIdx.y], acc);
}

#endif /* X265CU_WP_CUH */
