/* x265cu_me.cuh -- full-resolution motion search of one PU per warp for sm_100a.
 *
 * MotionEstimate::motionEstimate with ref->isLowres == false (encoder/motion.cpp:571-1172): the clipped predictor and the
 * quarter-pel predictor candidates (:597-641), the five integer patterns -- DIA :650-668, HEX :670-742, UMH :744-927, STAR
 * :929-1037 with StarPatternSearch :329-569, FULL :1039-1071 -- and the sub-pel ladder (:1085-1168, workload[] :45-55) on
 * subpelCompare (:1174-1203, luma: the lookahead's setSourcePU has no chroma) whose fractional blocks come from the 8-tap
 * filters of common/ipfilter.cpp (:80-119 horizontal, :166-205 vertical, :366-372 both through 14-bit intermediates).
 *
 * Mapping.  A search is a chain of decisions over small sets of candidate vectors whose costs do not depend on each other,
 * so the warp runs the scalar decision machine redundantly in every lane (warp-uniform control flow, no divergence) and
 * measures each SET cooperatively: the lanes split into groups of 32 / 16 / 8 / 4 / 2 per candidate (1, 2, 4, 8, 16
 * candidates at a time), a group strides over the PU's 4-sample words, SAD with __vsadu4 / __vsadu2 against the PU kept in
 * shared memory, a butterfly reduction inside the group.  The winner is then taken by scanning the costs in the reference's
 * order with its strict "<" -- the same first-minimum rule as its COPYn_IF_LT chains.  Sub-pel candidates are interpolated by
 * the whole warp into shared memory and measured there (SAD, or SATD as 4x4 Hadamard tiles: x265's satd of every PU shape is
 * the sum over 4x4 tiles of the halved abs-sum, pixel.cpp:192-242,979-1003).
 */
#ifndef X265CU_ME_CUH
#define X265CU_ME_CUH

struct MeItemDev                   /* = x265cu_me_item (include/x265cu.h) */
{
    int64_t offset;
    int16_t mvmin[2], mvmax[2], qmvp[2];
    int16_t numCandidates, merange;
    int16_t mvc[12][2];
};
struct MeResultDev { int16_t mv[2]; int32_t cost; };

/* a full-pel candidate; tag = shift that makes its mvcost argument (2; STAR's raster: 3 for every fourth vector) | STAR's point
 * number << 4 | STAR's distance << 8 */
struct MePt { short x, y; int tag; };

/* g_lumaFilter (common/constants.cpp:239-245) */
__constant__ short c_meLumaFilter[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 }, { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };
__constant__ signed char c_meHex2[8][2] = { { -1, -2 }, { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 }, { -2, 0 } };
__constant__ signed char c_meSquare1[9][2] = { { 0, 0 }, { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 }, { -1, -1 }, { -1, 1 }, { 1, -1 }, { 1, 1 } };
__constant__ signed char c_meHex4[16][2] = { { 0, -4 }, { 0, 4 }, { -2, -3 }, { 2, -3 }, { -4, -2 }, { 4, -2 }, { -4, -1 }, { 4, -1 },
                                             { -4, 0 }, { 4, 0 }, { -4, 1 }, { 4, 1 }, { -4, 2 }, { 4, 2 }, { -2, 3 }, { 2, 3 } };
__constant__ signed char c_meTwoPoint[16][2] = { { -1, 0 }, { 0, -1 }, { -1, -1 }, { 1, -1 }, { -1, 0 }, { 1, 0 }, { -1, 1 }, { -1, -1 },
                                                 { 1, -1 }, { 1, 1 }, { -1, 0 }, { 0, 1 }, { -1, 1 }, { 1, 1 }, { 1, 0 }, { 0, 1 } };
/* the small fixed patterns of UMH (:744-788, :835-837), as 4-point sets in the reference's order */
__constant__ signed char c_meUmhSets[7][4][2] = {
    { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 } },          /* 0 DIA1_ITER */
    { { 0, -2 }, { -1, -1 }, { 1, -1 }, { -2, 0 } },       /* 1, 2: the octagon */
    { { 2, 0 }, { -1, 1 }, { 1, 1 }, { 0, 2 } },
    { { -1, -2 }, { 1, -2 }, { -2, -1 }, { 2, -1 } },      /* 3, 4: the outer ring */
    { { -2, 1 }, { 2, 1 }, { -1, 2 }, { 1, 2 } },
    { { -2, -2 }, { -2, 2 }, { 2, -2 }, { 2, 2 } },        /* 5 corners after the big cross */
    { { 0, 0 }, { 0, 0 }, { 0, 0 }, { 0, 0 } } };
__constant__ unsigned char c_meRangeMul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };
/* workload[] (:45-55): hpel iterations, hpel directions, qpel iterations, qpel directions, hpel uses SATD */
__constant__ unsigned char c_meWorkload[8][5] = { { 1, 4, 0, 4, 0 }, { 1, 4, 1, 4, 0 }, { 1, 4, 1, 4, 1 }, { 2, 4, 1, 4, 1 },
                                                  { 2, 4, 2, 4, 1 }, { 1, 8, 1, 8, 1 }, { 2, 8, 1, 8, 1 }, { 2, 8, 2, 8, 1 } };

/* 4 samples from an aligned shared-memory address */
template <typename P> struct MeSm;
template <> struct MeSm<uint8_t>
{
    static __device__ __forceinline__ Px<uint8_t>::Row4 load(const uint8_t* p) { Px<uint8_t>::Row4 r; r.v = *(const uint32_t*)p; return r; }
};
template <> struct MeSm<uint16_t>
{
    static __device__ __forceinline__ Px<uint16_t>::Row4 load(const uint16_t* p) { const uint2 w = *(const uint2*)p; Px<uint16_t>::Row4 r; r.lo = w.x; r.hi = w.y; return r; }
};

template <typename P>
struct PuSearch
{
    /* warp-uniform */
    int w, h, wq, sizeScale, depth;
    const P* fref; int64_t rs;
    const uint16_t* lut;
    int mvpx, mvpy, minx, miny, maxx, maxy;
    int bx, by, bcost;
    int lane;
    /* this warp's shared memory */
    P* sFenc; P* sPred; short* sMid; MePt* sPts; int* sCost;

    __device__ __forceinline__ int mvcost(int qx, int qy) const
    {
        return (int)(uint16_t)(__ldg(lut + ((int)(short)qx - mvpx)) + __ldg(lut + ((int)(short)qy - mvpy)));
    }
    __device__ __forceinline__ bool inRange(int x, int y) const { return x >= minx && x <= maxx && y >= miny && y <= maxy; }

    /* costs (SAD + mvcost) of sPts[0..n), n <= 16, into sCost[] */
    __device__ void measure(int n)
    {
        __syncwarp();
        const int lg = n <= 1 ? 5 : (n <= 2 ? 4 : (n <= 4 ? 3 : (n <= 8 ? 2 : 1)));
        const int g = 1 << lg, p = lane >> lg, j = lane & (g - 1);
        int acc = 0;
        MePt pt = { 0, 0, 2 };
        int sh = 2;
        if (p < n)
        {
            pt = sPts[p];
            sh = pt.tag & 15;
            const P* r = fref + pt.x + (int64_t)pt.y * rs;
            int row = 0, col = j;
            while (col >= wq) { col -= wq; row++; }
            while (row < h)
            {
                acc += Px<P>::sad(MeSm<P>::load(sFenc + row * w + col * 4), Px<P>::load(r + row * rs + col * 4));
                col += g;
                while (col >= wq) { col -= wq; row++; }
            }
        }
        for (int o = g >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL_MASK, acc, o);
        if (p < n && j == 0)
            sCost[p] = acc + mvcost((int)(short)(pt.x << sh), (int)(short)(pt.y << sh));
        __syncwarp();
    }
    /* candidate i of the next set; a set always starts at i == 0 (every lane has read the previous set's results by then) */
    __device__ __forceinline__ void put(int i, int x, int y, int tag = 2)
    {
        if (i == 0) __syncwarp();
        if (lane == 0) { MePt t; t.x = (short)x; t.y = (short)y; t.tag = tag; sPts[i] = t; }
    }
    /* COST_MV over sPts[0..n) in order: the first strict minimum below bcost becomes (bx, by); returns its index or -1 */
    __device__ int take(int n)
    {
        measure(n);
        int won = -1;
        for (int i = 0; i < n; i++)
        {
            const int c = sCost[i];
            if (c < bcost) { bcost = c; won = i; }
        }
        if (won >= 0) { bx = sPts[won].x; by = sPts[won].y; }
        return won;
    }
    __device__ void trySet(int ox, int oy, int set)
    {
        for (int k = 0; k < 4; k++) put(k, ox + c_meUmhSets[set][k][0], oy + c_meUmhSets[set][k][1]);
        take(4);
    }

    /* SAD / SATD of the PU against a block (global: any alignment; shared: aligned rows of pitch w) */
    template <bool SHARED>
    __device__ int blockCost(const P* b, int64_t pitch, int useSatd)
    {
        int acc = 0;
        if (!useSatd)
        {
            int row = 0, col = lane;
            while (col >= wq) { col -= wq; row++; }
            while (row < h)
            {
                const typename Px<P>::Row4 v = SHARED ? MeSm<P>::load(b + row * pitch + col * 4) : Px<P>::load(b + row * pitch + col * 4);
                acc += Px<P>::sad(MeSm<P>::load(sFenc + row * w + col * 4), v);
                col += 32;
                while (col >= wq) { col -= wq; row++; }
            }
        }
        else
        {
            int ty = 0, tx = lane;
            while (tx >= wq) { tx -= wq; ty++; }
            const int th = h >> 2;
            while (ty < th)
            {
                typename Px<P>::Row4 f[4], r[4];
#pragma unroll
                for (int i = 0; i < 4; i++)
                {
                    f[i] = MeSm<P>::load(sFenc + (ty * 4 + i) * w + tx * 4);
                    r[i] = SHARED ? MeSm<P>::load(b + (ty * 4 + i) * pitch + tx * 4) : Px<P>::load(b + (ty * 4 + i) * pitch + tx * 4);
                }
                acc += satd4x4_abs<P>(f, r) >> 1;
                tx += 32;
                while (tx >= wq) { tx -= wq; ty++; }
            }
        }
        return warp_sum(acc);
    }

    /* subpelCompare (:1174-1203), luma */
    __device__ int subpel(int qx, int qy, int useSatd)
    {
        qx = (short)qx; qy = (short)qy;
        const P* r = fref + (qx >> 2) + (int64_t)(qy >> 2) * rs;
        const int xf = qx & 3, yf = qy & 3;
        if (!(xf | yf))
            return blockCost<false>(r, rs, useSatd);
        const int maxVal = (1 << depth) - 1;
        __syncwarp();
        if (!yf || !xf)
        {
            const int f = yf ? yf : xf;
            const int64_t step = yf ? rs : 1;
            int row = 0, col = lane;
            while (col >= w) { col -= w; row++; }
            while (row < h)
            {
                const P* p = r + col + (int64_t)row * rs - 3 * step;
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)__ldg(p + k * step) * c_meLumaFilter[f][k];
                int v = (int)(short)((sum + 32) >> 6);
                v = v < 0 ? 0 : (v > maxVal ? maxVal : v);
                sPred[row * w + col] = (P)v;
                col += 32;
                while (col >= w) { col -= w; row++; }
            }
        }
        else
        {
            const int head = 14 - depth, sh1 = 6 - head, off1 = -(8192 << sh1);
            const int sh2 = 6 + head, off2 = (1 << (sh2 - 1)) + (8192 << 6);
            int row = 0, col = lane;
            while (col >= w) { col -= w; row++; }
            while (row < h + 7)
            {
                const P* p = r + col - 3 + (int64_t)(row - 3) * rs;
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)__ldg(p + k) * c_meLumaFilter[xf][k];
                sMid[row * w + col] = (short)((sum + off1) >> sh1);
                col += 32;
                while (col >= w) { col -= w; row++; }
            }
            __syncwarp();
            row = 0; col = lane;
            while (col >= w) { col -= w; row++; }
            while (row < h)
            {
                int sum = 0;
#pragma unroll
                for (int k = 0; k < 8; k++) sum += (int)sMid[(row + k) * w + col] * c_meLumaFilter[yf][k];
                int v = (int)(short)((sum + off2) >> sh2);
                v = v < 0 ? 0 : (v > maxVal ? maxVal : v);
                sPred[row * w + col] = (P)v;
                col += 32;
                while (col >= w) { col -= w; row++; }
            }
        }
        __syncwarp();
        const int c = blockCost<true>(sPred, w, useSatd);
        __syncwarp();
        return c;
    }

    /* HEX (:670-742): hexagon walk, then the 8-point square */
    __device__ void hexWalk(int merange)
    {
        for (int k = 0; k < 6; k++) put(k, bx + c_meHex2[k + 1][0], by + c_meHex2[k + 1][1]);
        measure(6);
        int dir = -1;
        for (int k = 0; k < 6; k++) { const int c = sCost[k]; if (c < bcost) { bcost = c; dir = k; } }
        if (dir >= 0)
        {
            bx += c_meHex2[dir + 1][0]; by += c_meHex2[dir + 1][1];
            for (int i = (merange >> 1) - 1; i > 0 && inRange(bx, by); i--)
            {
                for (int k = 0; k < 3; k++) put(k, bx + c_meHex2[dir + k][0], by + c_meHex2[dir + k][1]);
                measure(3);
                int step = -1;
                for (int k = 0; k < 3; k++) { const int c = sCost[k]; if (c < bcost) { bcost = c; step = k; } }
                if (step < 0) break;
                dir += step - 1;
                dir = (dir + 6) % 6;
                bx += c_meHex2[dir + 1][0]; by += c_meHex2[dir + 1][1];
            }
        }
        for (int k = 0; k < 8; k++) put(k, bx + c_meSquare1[k + 1][0], by + c_meSquare1[k + 1][1]);
        take(8);
    }

    /* CROSS (:303-327): every point is relative to (ox, oy), so the whole list is independent of what it finds */
    __device__ void cross(int ox, int oy, int start, int xmax, int ymax)
    {
        for (int axis = 0; axis < 2; axis++)
        {
            const int lim = axis ? ymax : xmax;
            const int o = axis ? oy : ox, mn = axis ? miny : minx, mx = axis ? maxy : maxx;
            int n = 0;
            int i = (short)start;
            const bool fast = lim <= min(mx - o, o - mn);
#define ME_CROSS_PUT(d) do { if (axis) put(n, ox, oy + (d)); else put(n, ox + (d), oy); if (++n == 16) { take(16); n = 0; } } while (0)
            if (fast)
                for (; i < lim - 2; i = (short)(i + 4)) { ME_CROSS_PUT(i); ME_CROSS_PUT(-i); ME_CROSS_PUT(i + 2); ME_CROSS_PUT(-i - 2); }
            for (; i < lim; i = (short)(i + 2))
            {
                if (o + i <= mx) ME_CROSS_PUT(i);
                if (o - i >= mn) ME_CROSS_PUT(-i);
            }
#undef ME_CROSS_PUT
            if (n) take(n);
        }
    }

    /* one ring of StarPatternSearch: points gathered in the reference's order, then COST_MV_PT_DIST over them */
    int starN;
    __device__ __forceinline__ void starPut(int x, int y, bool ringInside, bool okA, bool okB, int point, int dist)
    {
        if (ringInside || (okA && okB))
            put(starN++, x, y, 2 | (point << 4) | (dist << 8));
    }
    __device__ void starTake(int& bPoint, int& bDist)
    {
        if (!starN) return;
        const int won = take(starN);
        if (won >= 0) { const int tag = sPts[won].tag; bPoint = (tag >> 4) & 15; bDist = tag >> 8; }
        starN = 0;
    }
    __device__ void starPattern(int& bPoint, int& bDist, int earlyExitIters, int merange)
    {
        const int ox = bx, oy = by;
        int rounds = 0;
        for (int dist = 1; dist <= 8; dist <<= 1)
        {
            const int top = (short)(oy - dist), bottom = (short)(oy + dist), left = (short)(ox - dist), right = (short)(ox + dist);
            const int top2 = (short)(oy - (dist >> 1)), bottom2 = (short)(oy + (dist >> 1)), left2 = (short)(ox - (dist >> 1)), right2 = (short)(ox + (dist >> 1));
            const int saved = bcost;
            const bool in = top >= miny && left >= minx && right <= maxx && bottom <= maxy;
            starN = 0;
            starPut(ox, top, in, top >= miny, true, 2, dist);
            if (dist > 1)
            {
                starPut(left2, top2, in, top2 >= miny, left2 >= minx, 1, dist >> 1);
                starPut(right2, top2, in, top2 >= miny, right2 <= maxx, 3, dist >> 1);
            }
            starPut(left, oy, in, left >= minx, true, 4, dist);
            starPut(right, oy, in, right <= maxx, true, 5, dist);
            if (dist > 1)
            {
                starPut(left2, bottom2, in, bottom2 <= maxy, left2 >= minx, 6, dist >> 1);
                starPut(right2, bottom2, in, bottom2 <= maxy, right2 <= maxx, 8, dist >> 1);
            }
            starPut(ox, bottom, in, bottom <= maxy, true, 7, dist);
            starTake(bPoint, bDist);
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
        for (int dist = 16; dist <= (int)(short)merange; dist = (short)(dist << 1))
        {
            const int top = (short)(oy - dist), bottom = (short)(oy + dist), left = (short)(ox - dist), right = (short)(ox + dist);
            const int saved = bcost;
            const bool in = top >= miny && left >= minx && right <= maxx && bottom <= maxy;
            starN = 0;
            starPut(ox, top, in, top >= miny, true, 0, dist);
            starPut(left, oy, in, left >= minx, true, 0, dist);
            starPut(right, oy, in, right <= maxx, true, 0, dist);
            starPut(ox, bottom, in, bottom <= maxy, true, 0, dist);
            for (int index = 1; index < 4; index++)
            {
                const int yt = (short)(top + (dist >> 2) * index), yb = (short)(bottom - (dist >> 2) * index);
                const int xl = (short)(ox - (dist >> 2) * index), xr = (short)(ox + (dist >> 2) * index);
                starPut(xl, yt, in, yt >= miny, xl >= minx, 0, dist);
                starPut(xr, yt, in, yt >= miny, xr <= maxx, 0, dist);
                starPut(xl, yb, in, yb <= maxy, xl >= minx, 0, dist);
                starPut(xr, yb, in, yb <= maxy, xr <= maxx, 0, dist);
            }
            starTake(bPoint, bDist);
            if (bcost < saved) rounds = 0;
            else if (++rounds >= earlyExitIters) return;
        }
    }
    __device__ void twoPoints(int point)
    {
        const int x0 = bx, y0 = by;
        int n = 0;
        for (int k = 0; k < 2; k++)
        {
            const int x = (short)(x0 + c_meTwoPoint[(point - 1) * 2 + k][0]), y = (short)(y0 + c_meTwoPoint[(point - 1) * 2 + k][1]);
            if (inRange(x, y)) put(n++, x, y);
        }
        if (n) take(n);
    }
    /* raster scans: STAR's refinement (step 5, `<< 3` in the fourth vector's mvcost, :971-1003) and FULL (step 1, :1039-1071).
     * Positions do not depend on results, so they are measured 16 at a time and scanned in order. */
    __device__ void raster(int step, int fourthShift)
    {
        int n = 0;
        for (int ty = miny; ty <= maxy; ty = (short)(ty + step))
            for (int tx = minx; tx <= maxx; tx = (short)(tx + step))
            {
                if (tx + step * 3 <= maxx)
                {
                    for (int k = 0; k < 4; k++)
                    {
                        if (k) tx = (short)(tx + step);
                        put(n, tx, ty, k == 3 ? fourthShift : 2);
                        if (++n == 16) { take(16); n = 0; }
                    }
                }
                else
                {
                    put(n, tx, ty);
                    if (++n == 16) { take(16); n = 0; }
                }
            }
        if (n) take(n);
    }

    /* UMH up to its final hexagon walk (:744-927): false = the pattern ended early */
    __device__ bool umh(int pmx, int pmy, int qmvpx, int qmvpy, int numCandidates, const int16_t (*mvc)[2], int& merange, bool is64)
    {
        int crossStart = 1;
        const int ucost1 = bcost;
        trySet(pmx, pmy, 0);
        if (pmx | pmy) trySet(0, 0, 0);
        const int ucost2 = bcost;
        if ((bx | by) && !(bx == pmx && by == pmy)) trySet(bx, by, 0);
        if (bcost == ucost2) crossStart = 3;
        int ox = bx, oy = by;
#define ME_THRESH(v) (bcost < (((v) >> 4) * sizeScale))
        if (bcost == ucost2 && ME_THRESH(2000))
        {
            trySet(ox, oy, 1);
            trySet(ox, oy, 2);
            if (bcost == ucost1 && ME_THRESH(500)) return false;
            if (bcost == ucost2)
            {
                const int range = (short)((short)(merange >> 1) | 1);
                cross(ox, oy, 3, range, range);
                trySet(ox, oy, 3);
                trySet(ox, oy, 4);
                if (bcost == ucost2) return false;
                crossStart = range + 2;
            }
        }
        if (numCandidates)
        {
            int mvd, denom = 1;
            if (numCandidates == 1)
                mvd = is64 ? 25 : abs(qmvpx - mvc[0][0]) + abs(qmvpy - mvc[0][1]);
            else
            {
                denom = numCandidates - 1;
                mvd = 0;
                if (!is64) { mvd = abs(qmvpx - mvc[0][0]) + abs(qmvpy - mvc[0][1]); denom++; }
                for (int i = 0; i < numCandidates - 1; i++)
                    mvd += abs(mvc[i][0] - mvc[i + 1][0]) + abs(mvc[i][1] - mvc[i + 1][1]);
            }
            const int sadCtx = ME_THRESH(1000) ? 0 : ME_THRESH(2000) ? 1 : ME_THRESH(4000) ? 2 : 3;
            const int mvdCtx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
            merange = (merange * c_meRangeMul[mvdCtx][sadCtx]) >> 2;
        }
#undef ME_THRESH
        cross(ox, oy, crossStart, merange, merange >> 1);
        trySet(ox, oy, 5);
        /* 16-point hexagon grid scaled 1 .. merange / 4 around the best vector so far */
        ox = bx; oy = by;
        unsigned i = 1;
        do
        {
            const bool slow = (int)(4 * i) > min(min(maxx - ox, ox - minx), min(maxy - oy, oy - miny));
            int n = 0;
            for (int j = 0; j < 16; j++)
            {
                const int x = (short)(ox + (short)(c_meHex4[j][0] * (short)i)), y = (short)(oy + (short)(c_meHex4[j][1] * (short)i));
                if (!slow || inRange(x, y)) put(n++, x, y);
            }
            if (n) take(n);
            i = (i + 1) & 0xffff;
        }
        while ((int)i <= (merange >> 2));
        return inRange(bx, by);
    }
};

/* one warp per search; dynamic shared memory = warps * (2 * w * h * sizeof(P) + (h + 7) * w * 2 + 16 * (8 + 4)) */
template <typename P>
__global__ void __launch_bounds__(128) pu_motion_search_kernel(int method, int subme, int w, int h, int depth, const P* __restrict__ fencPlane, int64_t fencStride,
                                                               const P* __restrict__ refPlane, int64_t refStride, const uint16_t* __restrict__ lutCentre,
                                                               int n, const MeItemDev* __restrict__ items, MeResultDev* __restrict__ out)
{
    extern __shared__ __align__(16) unsigned char meSmem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nWarps = blockDim.x >> 5;
    const int idx = blockIdx.x * nWarps + warp;
    if (idx >= n) return;
    const size_t blockBytes = (size_t)w * h * sizeof(P);
    const size_t midBytes = ((size_t)(h + 7) * w * 2 + 15) & ~(size_t)15;
    const size_t perWarp = 2 * blockBytes + midBytes + 16 * sizeof(MePt) + 16 * sizeof(int);
    unsigned char* base = meSmem + warp * perWarp;
    PuSearch<P> s;
    s.lane = lane;
    s.sFenc = (P*)base; s.sPred = (P*)(base + blockBytes); s.sMid = (short*)(base + 2 * blockBytes);
    s.sPts = (MePt*)(base + 2 * blockBytes + midBytes); s.sCost = (int*)(s.sPts + 16);
    s.w = w; s.h = h; s.wq = w >> 2; s.sizeScale = (h * h) >> 4; s.depth = depth;
    const MeItemDev& it = items[idx];
    s.fref = refPlane + it.offset; s.rs = refStride;
    s.lut = lutCentre;
    s.mvpx = it.qmvp[0]; s.mvpy = it.qmvp[1];
    s.minx = it.mvmin[0]; s.miny = it.mvmin[1]; s.maxx = it.mvmax[0]; s.maxy = it.mvmax[1];
    /* the PU into shared memory (setSourcePU's copy_pp into the FENC_STRIDE cache) */
    {
        const P* f = fencPlane + it.offset;
        int row = 0, col = lane;
        while (col >= w) { col -= w; row++; }
        while (row < h)
        {
            s.sFenc[row * w + col] = __ldg(f + (int64_t)row * fencStride + col);
            col += 32;
            while (col >= w) { col -= w; row++; }
        }
        __syncwarp();
    }
    const int qminx = (short)(s.minx << 2), qminy = (short)(s.miny << 2), qmaxx = (short)(s.maxx << 2), qmaxy = (short)(s.maxy << 2);
    int pqx = s.mvpx > qmaxx ? qmaxx : s.mvpx; if (pqx < qminx) pqx = qminx;
    int pqy = s.mvpy > qmaxy ? qmaxy : s.mvpy; if (pqy < qminy) pqy = qminy;
    int bestPreX = pqx, bestPreY = pqy;
    int bprecost = s.subpel(pqx, pqy, 0);
    const int pmx = (short)((pqx + 2) >> 2), pmy = (short)((pqy + 2) >> 2);
    s.bx = pmx; s.by = pmy; s.bcost = bprecost;
    {
        /* the rounded predictor (when the predictor was sub-pel) and the zero vector: one measured set, taken in order */
        int nn = 0, iRound = -1, iZero = -1;
        if ((pqx | pqy) & 3) { s.put(nn, pmx, pmy); iRound = nn++; }
        if (pqx | pqy) { s.put(nn, 0, 0); iZero = nn++; }
        if (nn)
        {
            s.measure(nn);
            if (iRound >= 0) s.bcost = s.sCost[iRound];
            if (iZero >= 0 && s.sCost[iZero] < s.bcost) { s.bcost = s.sCost[iZero]; s.bx = s.by = 0; }
        }
    }
    const int numCandidates = it.numCandidates;
    for (int i = 0; i < numCandidates; i++)
    {
        int mx = it.mvc[i][0] > qmaxx ? qmaxx : it.mvc[i][0]; if (mx < qminx) mx = qminx;
        int my = it.mvc[i][1] > qmaxy ? qmaxy : it.mvc[i][1]; if (my < qminy) my = qminy;
        if ((mx | my) && !(mx == pqx && my == pqy) && !(mx == bestPreX && my == bestPreY))
        {
            const int c = s.subpel(mx, my, 0) + s.mvcost(mx, my);
            if (c < bprecost) { bprecost = c; bestPreX = mx; bestPreY = my; }
        }
    }
    int merange = it.merange;
    bool walk = false;
    if (method == 0)
    {
        /* DIA (:650-668) */
        int i = merange;
        do
        {
            for (int k = 0; k < 4; k++) s.put(k, s.bx + c_meUmhSets[0][k][0], s.by + c_meUmhSets[0][k][1]);
            if (s.take(4) < 0) break;
        }
        while (--i && s.inRange(s.bx, s.by));
    }
    else if (method == 1)
        walk = true;
    else if (method == 2)
        walk = s.umh(pmx, pmy, s.mvpx, s.mvpy, numCandidates, it.mvc, merange, w == 64 && h == 64);
    else if (method == 3)
    {
        int bPoint = 0, bDist = 0;
        bool done = false;
        s.starPattern(bPoint, bDist, 3, merange);
        if (bDist == 1)
        {
            if (!bPoint) done = true;
            else
            {
                const int saved = s.bcost;
                s.twoPoints(bPoint);
                if (s.bcost == saved) done = true;
            }
        }
        if (!done)
        {
            if (bDist > 5) s.raster(5, 3);
            while (bDist > 0)
            {
                bDist = 0; bPoint = 0;
                s.starPattern(bPoint, bDist, 32, merange);
                if (bDist == 1)
                {
                    if (bPoint) s.twoPoints(bPoint);
                    break;
                }
            }
        }
    }
    else
        s.raster(1, 2);
    if (walk) s.hexWalk(merange);

    int qx, qy, bcost;
    if (bprecost < s.bcost) { qx = bestPreX; qy = bestPreY; bcost = bprecost; }
    else { qx = (short)(s.bx << 2); qy = (short)(s.by << 2); bcost = s.bcost; }
    if (!bcost)
        bcost = s.mvcost(qx, qy);
    else
    {
        const int hs = c_meWorkload[subme][4];
        if (hs) bcost = s.subpel(qx, qy, 1) + s.mvcost(qx, qy);
        for (int pass = 0; pass < 2; pass++)
        {
            const int iters = c_meWorkload[subme][pass ? 2 : 0], dirs = c_meWorkload[subme][pass ? 3 : 1];
            const int mul = pass ? 1 : 2, satd = pass ? 1 : hs;
            if (pass && !hs) bcost = s.subpel(qx, qy, 1) + s.mvcost(qx, qy);
            for (int iter = 0; iter < iters; iter++)
            {
                int bdir = 0;
                for (int i = 1; i <= dirs; i++)
                {
                    const int cx = (short)(qx + c_meSquare1[i][0] * mul), cy = (short)(qy + c_meSquare1[i][1] * mul);
                    const int c = s.subpel(cx, cy, satd) + s.mvcost(cx, cy);
                    if (c < bcost) { bcost = c; bdir = i; }
                }
                if (!bdir) break;
                qx = (short)(qx + c_meSquare1[bdir][0] * mul); qy = (short)(qy + c_meSquare1[bdir][1] * mul);
            }
        }
    }
    if (lane == 0)
    {
        out[idx].mv[0] = (int16_t)qx; out[idx].mv[1] = (int16_t)qy; out[idx].cost = bcost;
    }
}

#endif /* X265CU_ME_CUH */
