/* x265la_oracle.c -- CPU oracle (plain C restatement) of the x265 1.9 lookahead cost path.
 *
 * TEST INFRASTRUCTURE ONLY -- see x265la_oracle.h.  Every function cites the reference file:line
 * (relative to /root/reference/x265_1.9/source) whose behaviour it restates.  Written for
 * clarity, not speed: straightforward integer loops, no SWAR, no SIMD.
 */
#include "x265la_oracle.h"
#include "synth.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>

#define PIXEL_MAX ((1 << ORACLE_DEPTH) - 1)
#define CU 8
#define LOWRES_COST_MASK ((1 << 14) - 1)
#define LOWRES_COST_SHIFT 14

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }
static inline int iclip(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
static inline pixel clip_pixel(int v) { return (pixel)iclip(0, PIXEL_MAX, v); }

int ola_depth(void) { return ORACLE_DEPTH; }

/* ------------------------------------------------------------------------------------------
 * CRC32 (zlib polynomial) so Python can cross-check with zlib.crc32
 * ---------------------------------------------------------------------------------------- */
uint32_t ola_crc32(const void* p, size_t n)
{
    static uint32_t tab[256];
    if (!tab[1])
        for (uint32_t i = 0; i < 256; i++)
        {
            uint32_t c = i;
            for (int k = 0; k < 8; k++) c = (c & 1) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            tab[i] = c;
        }
    const uint8_t* b = (const uint8_t*)p;
    uint32_t crc = 0xFFFFFFFFu;
    while (n--) crc = tab[(crc ^ *b++) & 255] ^ (crc >> 8);
    return ~crc;
}

void ola_synth_frame(int w, int h, int t, int nframes, uint32_t seed, void* y, int ystride, void* u, void* v, int cstride)
{
    synth_frame(w, h, t, nframes, seed, ORACLE_DEPTH, y, ystride, u, v, cstride);
}

/* PicYuv::copyFromPicture padding (common/picyuv.cpp:168-178,287-298): the picture is extended to
 * a multiple of 16 plus ONE extra column/row by replicating the last column/row; the lowres
 * downscale reads that extra column/row.  dst must have room for (w + padx) x (h + pady). */
void ola_copy_picture(const pixel* src, int w, int h, pixel* dst, intptr_t dstStride)
{
    int padx = (w & 15) ? 16 - (w & 15) : 0;
    int pady = (h & 15) ? 16 - (h & 15) : 0;
    padx++; pady++;
    for (int y = 0; y < h; y++)
    {
        memcpy(dst + y * dstStride, src + (intptr_t)y * w, w * sizeof(pixel));
        for (int x = 0; x < padx; x++)
            dst[y * dstStride + w + x] = dst[y * dstStride + w - 1];
    }
    for (int i = 1; i <= pady; i++)
        memcpy(dst + (h - 1 + i) * dstStride, dst + (h - 1) * dstStride, (w + padx) * sizeof(pixel));
}

/* ------------------------------------------------------------------------------------------
 * Geometry: Lowres::create, common/lowres.cpp:30-48; Lookahead ctor, encoder/slicetype.cpp:503-506
 * ---------------------------------------------------------------------------------------- */
void ola_geometry(int srcW, int srcH, int marginX, int marginY, ola_geom* g)
{
    g->srcW = srcW; g->srcH = srcH;
    int w = srcW / 2, l = srcH / 2;
    g->stride = w + 2 * marginX;
    if (g->stride & 31) g->stride += 32 - (g->stride & 31);
    g->wCU = (w + CU - 1) >> 3;
    g->hCU = (l + CU - 1) >> 3;
    g->nCU = g->wCU * g->hCU;
    g->width = g->wCU * CU;
    g->lines = g->hCU * CU;
    g->marginX = marginX; g->marginY = marginY;
    g->paddedLines = g->lines + 2 * marginY;
    g->planeSize = (int64_t)g->stride * g->paddedLines;
    g->padOffset = (int64_t)g->stride * marginY + marginX;
}

/* cooperative slice geometry: Lookahead ctor, encoder/slicetype.cpp:534-558 */
void ola_coop_slices(int srcH, int lookaheadSlices, int hasPool, int hCU, int* numCoopSlices, int* numRowsPerSlice)
{
    if (!hasPool || srcH < 720) lookaheadSlices = 0;
    if (lookaheadSlices > 1)
    {
        int rps = hCU / lookaheadSlices;
        rps = imax(rps, 10);
        rps = imin(rps, hCU);
        *numRowsPerSlice = rps;
        *numCoopSlices = hCU / rps;
    }
    else
    {
        *numRowsPerSlice = hCU;
        *numCoopSlices = 1;
    }
}

ola_frame* ola_frame_create(int srcW, int srcH, int marginX, int marginY, int bframes, int aq)
{
    ola_frame* f = (ola_frame*)calloc(1, sizeof(ola_frame));
    ola_geometry(srcW, srcH, marginX, marginY, &f->g);
    f->bframes = bframes;
    f->hasAq = aq;
    int n = f->g.nCU;
    /* CHECKED_MALLOC_ZERO(buffer[0], pixel, 4 * planesize), lowres.cpp:62 */
    f->buffer[0] = (pixel*)calloc((size_t)(4 * f->g.planeSize), sizeof(pixel));
    for (int i = 0; i < 4; i++)
    {
        f->buffer[i] = f->buffer[0] + i * f->g.planeSize;
        f->plane[i] = f->buffer[i] + f->g.padOffset;
    }
    f->intraCost = (int32_t*)calloc(n, sizeof(int32_t));
    f->intraMode = (uint8_t*)calloc(n, 1);
    f->propagateCost = (uint16_t*)calloc(n, sizeof(uint16_t));
    if (aq)
    {
        f->invQscale = (int32_t*)calloc(n, sizeof(int32_t));
        f->qpAqOffset = (double*)calloc(n, sizeof(double));
        f->qpCuTreeOffset = (double*)calloc(n, sizeof(double));
        f->blockVariance = (uint32_t*)calloc(n, sizeof(uint32_t));
    }
    for (int i = 0; i < bframes + 2; i++)
        for (int j = 0; j < bframes + 2; j++)
        {
            f->rowSatds[i][j] = (int32_t*)calloc(f->g.hCU, sizeof(int32_t));
            f->lowresCosts[i][j] = (uint16_t*)calloc(n, sizeof(uint16_t));
        }
    for (int i = 0; i < bframes + 1; i++)
        for (int l = 0; l < 2; l++)
        {
            f->mvs[l][i] = (ola_mv*)calloc(n, sizeof(ola_mv));
            f->mvCosts[l][i] = (int32_t*)calloc(n, sizeof(int32_t));
        }
    return f;
}

void ola_frame_destroy(ola_frame* f)
{
    if (!f) return;
    free(f->buffer[0]); free(f->intraCost); free(f->intraMode); free(f->propagateCost);
    free(f->invQscale); free(f->qpAqOffset); free(f->qpCuTreeOffset); free(f->blockVariance);
    for (int i = 0; i < f->bframes + 2; i++)
        for (int j = 0; j < f->bframes + 2; j++) { free(f->rowSatds[i][j]); free(f->lowresCosts[i][j]); }
    for (int i = 0; i < f->bframes + 1; i++)
        for (int l = 0; l < 2; l++) { free(f->mvs[l][i]); free(f->mvCosts[l][i]); }
    free(f);
}

/* ------------------------------------------------------------------------------------------
 * mvcost LUT: BitCost::CalculateLogs + BitCost::setQP, encoder/bitcost.cpp:30-59,73-90.
 * lambda = x265_lambda_tab[X265_LOOKAHEAD_QP] (common/constants.cpp:74-125, common.h:208):
 * QP 12 -> 1.0 at 8 bit, QP 24 -> 16.0 at 10 bit, i.e. 2^((12-12)/6) * 2^(depth-8)... both are
 * exact powers of two in the reference table.
 * ---------------------------------------------------------------------------------------- */
static double ola_lambda(void)
{
#if ORACLE_DEPTH == 8
    return 1.0;
#elif ORACLE_DEPTH == 10
    return 16.0;
#else
    return 64.0;
#endif
}

int ola_lambda_int(void) { return (int)ola_lambda(); }

void ola_mvcost_table(uint16_t* out)
{
    const int M = 2 * 32768;
    double lambda = ola_lambda();
    float log2_2 = 2.0f / logf(2.0f);
    for (int i = 0; i <= M; i++)
    {
        float bits = i ? logf((float)(i + 1)) * log2_2 + 1.718f : 0.718f;
        double v = bits * lambda + 0.5f;
        if (v > 32767.0) v = 32767.0;
        out[M + i] = out[M - i] = (uint16_t)v;
    }
}

ola_ctx* ola_ctx_create(int bFrameBias, int numCoopSlices, int numRowsPerSlice)
{
    ola_ctx* c = (ola_ctx*)calloc(1, sizeof(ola_ctx));
    c->mvcostBase = (uint16_t*)malloc((4 * 32768 + 1) * sizeof(uint16_t));
    ola_mvcost_table(c->mvcostBase);
    c->mvcost = c->mvcostBase + 2 * 32768;
    c->lambda = ola_lambda_int();
    c->bFrameBias = bFrameBias;
    c->numCoopSlices = numCoopSlices;
    c->numRowsPerSlice = numRowsPerSlice;
    return c;
}

void ola_ctx_destroy(ola_ctx* c)
{
    if (!c) return;
    free(c->mvcostBase);
    free(c->wbuffer[0]);
    free(c);
}

/* ------------------------------------------------------------------------------------------
 * Pixel primitives, common/pixel.cpp
 * ---------------------------------------------------------------------------------------- */

/* sad<8,8>, pixel.cpp:39-54 */
int ola_sad8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int sum = 0;
    for (int y = 0; y < 8; y++, a += sa, b += sb)
        for (int x = 0; x < 8; x++)
            sum += abs((int)a[x] - (int)b[x]);
    return sum;
}

/* 4x4 Hadamard of a difference block, sum of absolute transformed coefficients (not halved) */
static int hadamard4x4_abs(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int d[4][4], t[4][4];
    for (int y = 0; y < 4; y++)
        for (int x = 0; x < 4; x++)
            d[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    for (int y = 0; y < 4; y++)
    {
        int s01 = d[y][0] + d[y][1], d01 = d[y][0] - d[y][1];
        int s23 = d[y][2] + d[y][3], d23 = d[y][2] - d[y][3];
        t[y][0] = s01 + s23; t[y][1] = d01 + d23; t[y][2] = s01 - s23; t[y][3] = d01 - d23;
    }
    int sum = 0;
    for (int x = 0; x < 4; x++)
    {
        int s01 = t[0][x] + t[1][x], d01 = t[0][x] - t[1][x];
        int s23 = t[2][x] + t[3][x], d23 = t[2][x] - t[3][x];
        sum += abs(s01 + s23) + abs(d01 + d23) + abs(s01 - s23) + abs(d01 - d23);
    }
    return sum;
}

/* satd_4x4, pixel.cpp:163-189: (sum of |H4 D H4|) >> 1 */
int ola_satd4x4(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    return hadamard4x4_abs(a, sa, b, sb) >> 1;
}

/* satd8<8,8> = satd_8x4(top) + satd_8x4(bottom), pixel.cpp:192-242,981; each satd_8x4 is the sum
 * of two 4x4 Hadamard abs-sums, halved once (the SWAR lanes of the reference never overflow for
 * legal pixels, SURVEY.md §8 a9, so plain int arithmetic is bit-equal). */
int ola_satd8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int total = 0;
    for (int row = 0; row < 8; row += 4)
    {
        int s = hadamard4x4_abs(a + row * sa, sa, b + row * sb, sb) +
                hadamard4x4_abs(a + row * sa + 4, sa, b + row * sb + 4, sb);
        total += s >> 1;
    }
    return total;
}

/* _sa8d_8x8, pixel.cpp:244-283: un-normalised sum of |H8 D H8| */
static int sa8d_raw(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int m[8][8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            m[y][x] = (int)a[y * sa + x] - (int)b[y * sb + x];
    /* 8-point Hadamard along rows then columns (butterfly order does not matter for the abs-sum) */
    for (int pass = 0; pass < 2; pass++)
        for (int i = 0; i < 8; i++)
        {
            int v[8];
            for (int k = 0; k < 8; k++) v[k] = pass ? m[k][i] : m[i][k];
            for (int step = 1; step < 8; step <<= 1)
                for (int k = 0; k < 8; k += step << 1)
                    for (int j = k; j < k + step; j++)
                    {
                        int p = v[j], q = v[j + step];
                        v[j] = p + q; v[j + step] = p - q;
                    }
            for (int k = 0; k < 8; k++) { if (pass) m[k][i] = v[k]; else m[i][k] = v[k]; }
        }
    int sum = 0;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            sum += abs(m[y][x]);
    return sum;
}

/* sa8d_8x8, pixel.cpp:285-288 */
int ola_sa8d8x8(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    return (sa8d_raw(a, sa, b, sb) + 2) >> 2;
}

/* sa8d_16x16, pixel.cpp:290-300: four raw 8x8 sums, rounded once */
int ola_sa8d16x16(const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int sum = sa8d_raw(a, sa, b, sb) + sa8d_raw(a + 8, sa, b + 8, sb) +
              sa8d_raw(a + 8 * sa, sa, b + 8 * sb, sb) + sa8d_raw(a + 8 + 8 * sa, sa, b + 8 + 8 * sb, sb);
    return (sum + 2) >> 2;
}

/* pixelavg_pp<8,8>, pixel.cpp:490-502 (weight argument ignored by the reference) */
void ola_pixelavg8x8(pixel* dst, intptr_t ds, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            dst[y * ds + x] = (pixel)(((int)a[y * sa + x] + (int)b[y * sb + x] + 1) >> 1);
}

/* frame_init_lowres_core, pixel.cpp:549-573 */
static inline int lowres_filter(int a, int b, int c, int d)
{
    return (((a + b + 1) >> 1) + ((c + d + 1) >> 1) + 1) >> 1;
}

void ola_frame_init_lowres(const pixel* src, pixel* d0, pixel* dh, pixel* dv, pixel* dc, intptr_t ss, intptr_t ds, int w, int h)
{
    for (int y = 0; y < h; y++)
    {
        const pixel* r0 = src + (intptr_t)(2 * y) * ss;
        const pixel* r1 = r0 + ss;
        const pixel* r2 = r1 + ss;
        for (int x = 0; x < w; x++)
        {
            int c0 = 2 * x, c1 = 2 * x + 1, c2 = 2 * x + 2;
            d0[y * ds + x] = (pixel)lowres_filter(r0[c0], r1[c0], r0[c1], r1[c1]);
            dh[y * ds + x] = (pixel)lowres_filter(r0[c1], r1[c1], r0[c2], r1[c2]);
            dv[y * ds + x] = (pixel)lowres_filter(r1[c0], r2[c0], r1[c1], r2[c1]);
            dc[y * ds + x] = (pixel)lowres_filter(r1[c1], r2[c1], r1[c2], r2[c2]);
        }
    }
}

/* extendPicBorder, pixel.cpp:908-922 + extendCURowColBorder, ipfilter.cpp:59-77.  Note the top and
 * bottom copies move a full `stride` of samples starting at column -marginX. */
void ola_extend_border(pixel* pic, intptr_t stride, int w, int h, int mx, int my)
{
    for (int y = 0; y < h; y++)
    {
        pixel* row = pic + y * stride;
        for (int x = 0; x < mx; x++)
        {
            row[-mx + x] = row[0];
            row[w + x] = row[w - 1];
        }
    }
    pixel* top = pic - mx;
    for (int y = 0; y < my; y++)
        memcpy(top - (y + 1) * stride, top, stride * sizeof(pixel));
    pixel* bot = pic - mx + (intptr_t)(h - 1) * stride;
    for (int y = 0; y < my; y++)
        memcpy(bot + (y + 1) * stride, bot, stride * sizeof(pixel));
}

/* weight_pp_c, pixel.cpp:463-488 */
void ola_weight_pp(const pixel* src, pixel* dst, intptr_t stride, int w, int h, int w0, int round, int shift, int offset)
{
    const int correction = 14 - ORACLE_DEPTH; /* IF_INTERNAL_PREC - X265_DEPTH */
    for (int y = 0; y < h; y++, src += stride, dst += stride)
        for (int x = 0; x < w; x++)
        {
            int16_t val = (int16_t)(src[x] << correction);
            dst[x] = clip_pixel(((w0 * val + round) >> shift) + offset);
        }
}

/* pixel_var<size>, pixel.cpp:649-666: sum in the low 32 bits, sum of squares in the high 32 */
static uint64_t var_n(const pixel* p, intptr_t s, int n)
{
    uint32_t sum = 0, sqr = 0;
    for (int y = 0; y < n; y++, p += s)
        for (int x = 0; x < n; x++)
        {
            sum += p[x];
            sqr += (uint32_t)p[x] * p[x];
        }
    return sum + ((uint64_t)sqr << 32);
}
uint64_t ola_var16(const pixel* p, intptr_t s) { return var_n(p, s, 16); }
uint64_t ola_var8(const pixel* p, intptr_t s) { return var_n(p, s, 8); }

/* x265_exp2fix8, common/common.cpp:94-101.  The 64-entry LUT (common/constants.cpp) is
 * round(256 * (2^(i/64) - 1)); regenerated here rather than copied. */
int ola_exp2fix8(double x)
{
    static uint8_t lut[64];
    static int init;
    if (!init)
    {
        for (int i = 0; i < 64; i++)
            lut[i] = (uint8_t)floor(256.0 * (pow(2.0, i / 64.0) - 1.0) + 0.5);
        init = 1;
    }
    int i = (int)(x * (-64.f / 6.f) + 512.5f);
    if (i < 0) return 0;
    if (i > 1023) return 0xffff;
    return (lut[i & 63] + 256) << (i >> 6) >> 8;
}

/* ------------------------------------------------------------------------------------------
 * Intra predictors at 8x8, common/intrapred.cpp
 * neighbour layout (33 samples): [0]=top-left, [1..16]=top + top-right, [17..32]=left + bottom-left
 * ---------------------------------------------------------------------------------------- */

/* intraFilter<8>, intrapred.cpp:31-51 */
void ola_intra_filter8(const pixel* s, pixel* f)
{
    const int n2 = 16;
    f[0] = (pixel)((2 * s[0] + s[1] + s[n2 + 1] + 2) >> 2);
    for (int i = 1; i < n2; i++)
        f[i] = (pixel)((2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2);
    f[n2] = s[n2];
    f[n2 + 1] = (pixel)((2 * s[n2 + 1] + s[0] + s[n2 + 2] + 2) >> 2);
    for (int i = n2 + 2; i < 2 * n2; i++)
        f[i] = (pixel)((2 * s[i] + s[i - 1] + s[i + 1] + 2) >> 2);
    f[2 * n2] = s[2 * n2];
}

/* intra_pred_dc_c<8> + dcPredFilter, intrapred.cpp:53-85 */
static void pred_dc8(pixel* dst, intptr_t ds, const pixel* s, int bFilter)
{
    int dc = 8;
    for (int i = 0; i < 8; i++)
        dc += s[1 + i] + s[17 + i];
    dc /= 16;
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            dst[y * ds + x] = (pixel)dc;
    if (bFilter)
    {
        const pixel* above = s + 1;
        const pixel* left = s + 17;
        dst[0] = (pixel)((above[0] + left[0] + 2 * dc + 2) >> 2);
        for (int x = 1; x < 8; x++)
            dst[x] = (pixel)((above[x] + 3 * dc + 2) >> 2);
        for (int y = 1; y < 8; y++)
            dst[y * ds] = (pixel)((left[y] + 3 * dc + 2) >> 2);
    }
}

/* planar_pred_c<3>, intrapred.cpp:87-100 */
static void pred_planar8(pixel* dst, intptr_t ds, const pixel* s)
{
    const pixel* above = s + 1;
    const pixel* left = s + 17;
    int topRight = above[8], bottomLeft = left[8];
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            dst[y * ds + x] = (pixel)(((7 - x) * left[y] + (7 - y) * above[x] + (x + 1) * topRight + (y + 1) * bottomLeft + 8) >> 4);
}

/* intra_pred_ang_c<8>, intrapred.cpp:102-204 */
static void pred_ang8(pixel* dst, intptr_t ds, const pixel* s0, int mode, int bFilter)
{
    static const int8_t angleTable[17] = { -32, -26, -21, -17, -13, -9, -5, -2, 0, 2, 5, 9, 13, 17, 21, 26, 32 };
    static const int16_t invAngleTable[8] = { 4096, 1638, 910, 630, 482, 390, 315, 256 };
    int hor = mode < 18;
    pixel nb[33];
    const pixel* s = s0;
    if (hor)
    {
        /* swap the roles of the top and left neighbours; the block is transposed at the end */
        nb[0] = s0[0];
        for (int i = 0; i < 16; i++) { nb[1 + i] = s0[17 + i]; nb[17 + i] = s0[1 + i]; }
        s = nb;
    }
    int angleOffset = hor ? 10 - mode : mode - 26;
    int angle = angleTable[8 + angleOffset];
    pixel blk[8][8];
    if (!angle)
    {
        for (int y = 0; y < 8; y++)
            for (int x = 0; x < 8; x++)
                blk[y][x] = s[1 + x];
        if (bFilter)
        {
            int topLeft = s[0], top = s[1];
            for (int y = 0; y < 8; y++)
                blk[y][0] = clip_pixel((int16_t)(top + ((s[17 + y] - topLeft) >> 1)));
        }
    }
    else
    {
        pixel refBuf[64];
        const pixel* ref;
        if (angle < 0)
        {
            int nbProjected = -((8 * angle) >> 5) - 1;
            pixel* rp = refBuf + nbProjected + 1;
            int invAngle = invAngleTable[-angleOffset - 1];
            int invAngleSum = 128;
            for (int i = 0; i < nbProjected; i++)
            {
                invAngleSum += invAngle;
                rp[-2 - i] = s[16 + (invAngleSum >> 8)];
            }
            for (int i = 0; i < 9; i++)
                rp[-1 + i] = s[i];
            ref = rp;
        }
        else
            ref = s + 1;
        int angleSum = 0;
        for (int y = 0; y < 8; y++)
        {
            angleSum += angle;
            int off = angleSum >> 5, frac = angleSum & 31;
            for (int x = 0; x < 8; x++)
                blk[y][x] = frac ? (pixel)(((32 - frac) * ref[off + x] + frac * ref[off + x + 1] + 16) >> 5) : ref[off + x];
        }
    }
    for (int y = 0; y < 8; y++)
        for (int x = 0; x < 8; x++)
            dst[y * ds + x] = hor ? blk[x][y] : blk[y][x];
}

void ola_intra_pred8(int mode, pixel* dst, intptr_t ds, const pixel* src, int bFilter)
{
    if (mode == 0) pred_planar8(dst, ds, src);
    else if (mode == 1) pred_dc8(dst, ds, src, bFilter);
    else pred_ang8(dst, ds, src, mode, bFilter);
}

/* g_intraFilterFlags[mode] & 8 (common/constants.cpp:550-556): at 8x8 only planar(0) and the
 * diagonal modes 2, 18, 34 use filtered neighbours. */
static inline int intra_filter_flag8(int mode) { return mode == 0 || mode == 2 || mode == 18 || mode == 34; }

/* ------------------------------------------------------------------------------------------
 * Lowres::init, common/lowres.cpp:128-165
 * ---------------------------------------------------------------------------------------- */
void ola_frame_init(ola_frame* f, const pixel* srcLuma, intptr_t srcStride, int poc)
{
    f->frameNum = poc;
    memset(f->costEst, -1, sizeof(f->costEst));
    memset(f->weightedCostDelta, 0, sizeof(f->weightedCostDelta));
    if (f->hasAq)
        memset(f->costEstAq, -1, sizeof(f->costEstAq));
    for (int y = 0; y < f->bframes + 2; y++)
        for (int x = 0; x < f->bframes + 2; x++)
            f->rowSatds[y][x][0] = -1;
    for (int i = 0; i < f->bframes + 1; i++)
    {
        f->mvs[0][i][0].x = OLA_MV_SENTINEL;
        f->mvs[1][i][0].x = OLA_MV_SENTINEL;
    }
    for (int i = 0; i < f->bframes + 2; i++)
        f->intraMbs[i] = 0;
    ola_frame_init_lowres(srcLuma, f->plane[0], f->plane[1], f->plane[2], f->plane[3], srcStride, f->g.stride, f->g.width, f->g.lines);
    for (int i = 0; i < 4; i++)
        ola_extend_border(f->plane[i], f->g.stride, f->g.width, f->g.lines, f->g.marginX, f->g.marginY);
}

/* ------------------------------------------------------------------------------------------
 * LookaheadTLD::calcAdaptiveQuantFrame + acEnergyCu, encoder/slicetype.cpp:48-228 (4:2:0)
 * y/u/v point at sample (0,0) of planes padded like PicYuv::copyFromPicture.
 * ---------------------------------------------------------------------------------------- */
static uint32_t ac_energy_plane(ola_frame* f, uint64_t sum_ssd, int shift, int plane)
{
    uint32_t sum = (uint32_t)sum_ssd;
    uint32_t ssd = (uint32_t)(sum_ssd >> 32);
    f->wp_sum[plane] += sum;
    f->wp_ssd[plane] += ssd;
    return ssd - (uint32_t)(((uint64_t)sum * sum) >> shift);
}

static uint32_t ac_energy_cu(ola_frame* f, const pixel* y, intptr_t ys, const pixel* u, const pixel* v, intptr_t cs, int bx, int by)
{
    uint32_t var = ac_energy_plane(f, ola_var16(y + bx + (intptr_t)by * ys, ys), 8, 0);
    if (u && v)
    {
        intptr_t co = (bx >> 1) + (intptr_t)(by >> 1) * cs;
        var += ac_energy_plane(f, ola_var8(u + co, cs), 6, 1);
        var += ac_energy_plane(f, ola_var8(v + co, cs), 6, 2);
    }
    return var;
}

void ola_aq_frame(ola_frame* f, const pixel* y, intptr_t ys, const pixel* u, const pixel* v, intptr_t cs,
                  int aqMode, double aqStrength, int weightp)
{
    int maxCol = f->g.srcW, maxRow = f->g.srcH;
    int blockCount = f->g.nCU;
    for (int i = 0; i < 3; i++) { f->wp_ssd[i] = 0; f->wp_sum[i] = 0; }
    int blockXY = 0;
    double strength = 0.f;
    if (aqMode == 0 || aqStrength == 0)
    {
        if (aqMode && aqStrength == 0)
        {
            for (int i = 0; i < blockCount; i++)
            {
                f->qpCuTreeOffset[i] = f->qpAqOffset[i] = 0;
                f->invQscale[i] = 256;
            }
        }
        if (weightp)
            for (int by = 0; by < maxRow; by += 16)
                for (int bx = 0; bx < maxCol; bx += 16)
                    ac_energy_cu(f, y, ys, u, v, cs, bx, by);
    }
    else
    {
        double avg_adj_pow2 = 0, avg_adj = 0, qp_adj = 0;
        double bias_strength = 0.f;
        if (aqMode == 2 || aqMode == 3)
        {
            double bit_depth_correction = 1.f / (1 << (2 * (ORACLE_DEPTH - 8)));
            f->frameVariance = 0;
            for (int by = 0; by < maxRow; by += 16)
            {
                uint64_t rowVariance = 0;
                for (int bx = 0; bx < maxCol; bx += 16)
                {
                    uint32_t energy = ac_energy_cu(f, y, ys, u, v, cs, bx, by);
                    f->blockVariance[blockXY] = energy;
                    rowVariance += energy;
                    qp_adj = pow(energy * bit_depth_correction + 1, 0.1);
                    f->qpCuTreeOffset[blockXY] = qp_adj;
                    avg_adj += qp_adj;
                    avg_adj_pow2 += qp_adj * qp_adj;
                    blockXY++;
                }
                f->frameVariance += (rowVariance / maxCol);
            }
            f->frameVariance /= maxRow;
            avg_adj /= blockCount;
            avg_adj_pow2 /= blockCount;
            strength = aqStrength * avg_adj;
            avg_adj = avg_adj - 0.5f * (avg_adj_pow2 - (11.f)) / avg_adj;
            bias_strength = aqStrength;
        }
        else
            strength = aqStrength * 1.0397f;

        blockXY = 0;
        for (int by = 0; by < maxRow; by += 16)
            for (int bx = 0; bx < maxCol; bx += 16)
            {
                if (aqMode == 3)
                {
                    qp_adj = f->qpCuTreeOffset[blockXY];
                    qp_adj = strength * (qp_adj - avg_adj) + bias_strength * (1.f - 11.f / (qp_adj * qp_adj));
                }
                else if (aqMode == 2)
                {
                    qp_adj = f->qpCuTreeOffset[blockXY];
                    qp_adj = strength * (qp_adj - avg_adj);
                }
                else
                {
                    uint32_t energy = ac_energy_cu(f, y, ys, u, v, cs, bx, by);
                    qp_adj = strength * (log2((double)(energy > 1 ? energy : 1)) - (14.427f + 2 * (ORACLE_DEPTH - 8)));
                }
                f->qpAqOffset[blockXY] = qp_adj;
                f->qpCuTreeOffset[blockXY] = qp_adj;
                f->invQscale[blockXY] = ola_exp2fix8(qp_adj);
                blockXY++;
            }
    }
    if (weightp)
    {
        maxCol = ((maxCol + 8) >> 4) << 4;
        maxRow = ((maxRow + 8) >> 4) << 4;
        int width[3] = { maxCol, maxCol >> 1, maxCol >> 1 };
        int height[3] = { maxRow, maxRow >> 1, maxRow >> 1 };
        for (int i = 0; i < 3; i++)
        {
            uint64_t sum = f->wp_sum[i], ssd = f->wp_ssd[i];
            f->wp_ssd[i] = ssd - (sum * sum + (width[i] * height[i]) / 2) / (width[i] * height[i]);
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * LookaheadTLD::lowresIntraEstimate, encoder/slicetype.cpp:230-336
 * ---------------------------------------------------------------------------------------- */
void ola_intra_estimate(ola_frame* f, int lambda)
{
    const ola_geom* g = &f->g;
    const int intraPenalty = 5 * lambda;
    const int lowresPenalty = 4;
    int64_t costEst = 0, costEstAq = 0;
    pixel pred[64], fenc[64], nb[33], nbf[33];

    for (int cuY = 0; cuY < g->hCU; cuY++)
    {
        f->rowSatds[0][0][cuY] = 0;
        for (int cuX = 0; cuX < g->wCU; cuX++)
        {
            int cuXY = cuX + cuY * g->wCU;
            const pixel* pix = f->plane[0] + CU * cuX + (intptr_t)CU * cuY * g->stride;
            for (int y = 0; y < 8; y++)
                memcpy(fenc + 8 * y, pix + (intptr_t)y * g->stride, 8 * sizeof(pixel));
            /* 17 samples of the row above starting at the top-left, then 16 of the column to the left */
            const pixel* tl = pix - g->stride - 1;
            memcpy(nb, tl, 17 * sizeof(pixel));
            for (int i = 1; i <= 16; i++)
                nb[16 + i] = tl[(intptr_t)i * g->stride];
            ola_intra_filter8(nb, nbf);

            int icost = OLA_COST_MAX, ilow = 0, cost;
            pred_dc8(pred, 8, nb, 1);
            cost = ola_satd8x8(fenc, 8, pred, 8);
            if (cost < icost) { icost = cost; ilow = 1; }
            pred_planar8(pred, 8, nbf);
            cost = ola_satd8x8(fenc, 8, pred, 8);
            if (cost < icost) { icost = cost; ilow = 0; }

            int acost = OLA_COST_MAX, alow = 4;
            for (int mode = 5; mode < 35; mode += 5)
            {
                pred_ang8(pred, 8, intra_filter_flag8(mode) ? nbf : nb, mode, 1);
                cost = ola_satd8x8(fenc, 8, pred, 8);
                if (cost < acost) { acost = cost; alow = mode; }
            }
            for (int dist = 2; dist >= 1; dist--)
            {
                int minusmode = alow - dist, plusmode = alow + dist;
                pred_ang8(pred, 8, intra_filter_flag8(minusmode) ? nbf : nb, minusmode, 1);
                cost = ola_satd8x8(fenc, 8, pred, 8);
                if (cost < acost) { acost = cost; alow = minusmode; }
                pred_ang8(pred, 8, intra_filter_flag8(plusmode) ? nbf : nb, plusmode, 1);
                cost = ola_satd8x8(fenc, 8, pred, 8);
                if (cost < acost) { acost = cost; alow = plusmode; }
            }
            if (acost < icost) { icost = acost; ilow = alow; }

            icost += intraPenalty + lowresPenalty;
            f->lowresCosts[0][0][cuXY] = (uint16_t)imin(icost, LOWRES_COST_MASK);
            f->intraCost[cuXY] = icost;
            f->intraMode[cuXY] = (uint8_t)ilow;

            int scored = (cuX > 0 && cuX < g->wCU - 1 && cuY > 0 && cuY < g->hCU - 1) || g->wCU <= 2 || g->hCU <= 2;
            int icostAq = (scored && f->invQscale) ? ((icost * f->invQscale[cuXY] + 128) >> 8) : icost;
            if (scored) { costEst += icost; costEstAq += icostAq; }
            f->rowSatds[0][0][cuY] += icostAq;
        }
    }
    f->costEst[0][0] = costEst;
    f->costEstAq[0][0] = costEstAq;
}

/* ------------------------------------------------------------------------------------------
 * Lowres motion compensation: ReferencePlanes::lowresMC / lowresQPelCost, common/lowres.h:62-103
 * ---------------------------------------------------------------------------------------- */
typedef struct ref_planes { pixel* plane[4]; intptr_t stride; } ref_planes;

/* public wrapper used by tests/core_emul.cpp */
static void lowres_mc(const ref_planes* r, intptr_t blockOffset, int qx, int qy, pixel* blk);
void ola_lowres_mc(pixel* const planes[4], intptr_t stride, intptr_t blockOffset, int qx, int qy, pixel* blk)
{
    ref_planes r = { { planes[0], planes[1], planes[2], planes[3] }, stride };
    lowres_mc(&r, blockOffset, qx, qy, blk);
}

/* fills blk[64] (stride 8) with the reference block for quarter-pel MV (qx,qy) */
static void lowres_mc(const ref_planes* r, intptr_t blockOffset, int qx, int qy, pixel* blk)
{
    int hpelA = (qy & 2) | ((qx & 2) >> 1);
    const pixel* a = r->plane[hpelA] + blockOffset + (qx >> 2) + (intptr_t)(qy >> 2) * r->stride;
    if ((qx | qy) & 1)
    {
        int qx2 = qx + (qx & 1), qy2 = qy + (qy & 1);
        int hpelB = (qy2 & 2) | ((qx2 & 2) >> 1);
        const pixel* b = r->plane[hpelB] + blockOffset + (qx2 >> 2) + (intptr_t)(qy2 >> 2) * r->stride;
        ola_pixelavg8x8(blk, 8, a, r->stride, b, r->stride);
    }
    else
        for (int y = 0; y < 8; y++)
            memcpy(blk + 8 * y, a + (intptr_t)y * r->stride, 8 * sizeof(pixel));
}

/* ------------------------------------------------------------------------------------------
 * MotionEstimate::motionEstimate, lowres / HEX / subpelRefine 1 branch,
 * encoder/motion.cpp:571-624,670-742,1081-1119 (+ setMVP/mvcost, bitcost.h:42-45)
 * ---------------------------------------------------------------------------------------- */
typedef struct me_state
{
    ola_ctx* c;
    const pixel* fenc;       /* 8x8, stride 8 */
    const ref_planes* ref;
    intptr_t blockOffset;
    int mvpx, mvpy;          /* unclipped MVP in qpel (setMVP) */
} me_state;

static inline int me_mvcost(const me_state* s, int qx, int qy)
{
    return (uint16_t)(s->c->mvcost[qx - s->mvpx] + s->c->mvcost[qy - s->mvpy]);
}

static int me_sad_fpel(me_state* s, int fx, int fy)
{
    const pixel* p = s->ref->plane[0] + s->blockOffset + fx + (intptr_t)fy * s->ref->stride;
    s->c->nSad++;
    return ola_sad8x8(s->fenc, 8, p, s->ref->stride);
}

static int me_qpel_cost(me_state* s, int qx, int qy, int useSatd)
{
    pixel blk[64];
    lowres_mc(s->ref, s->blockOffset, qx, qy, blk);
    if (useSatd) { s->c->nSatd++; return ola_satd8x8(s->fenc, 8, blk, 8); }
    s->c->nSad++;
    return ola_sad8x8(s->fenc, 8, blk, 8);
}

static int motion_estimate(me_state* s, int minx, int miny, int maxx, int maxy, int* outx, int* outy)
{
    static const int hex2[8][2] = { { -1, -2 }, { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 }, { -2, 0 } };
    static const int mod6m1[8] = { 5, 0, 1, 2, 3, 4, 5, 0 };
    static const int square1[9][2] = { { 0, 0 }, { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 }, { -1, -1 }, { -1, 1 }, { 1, -1 }, { 1, 1 } };
    const int merange = 16;

    /* clipped qpel MVP and its SAD (no mvcost) */
    int pmx = iclip(minx * 4, maxx * 4, s->mvpx), pmy = iclip(miny * 4, maxy * 4, s->mvpy);
    int bprecost = me_qpel_cost(s, pmx, pmy, 0);
    int bmx = (pmx + 2) >> 2, bmy = (pmy + 2) >> 2;
    int bcost = bprecost;
    if ((pmx | pmy) & 3)
        bcost = me_sad_fpel(s, bmx, bmy) + me_mvcost(s, bmx * 4, bmy * 4);
    if (pmx | pmy)
    {
        int cost = me_sad_fpel(s, 0, 0) + me_mvcost(s, 0, 0);
        if (cost < bcost) { bcost = cost; bmx = bmy = 0; }
    }

    /* hexagon, radius 2: six points in the order of hex2[1..6], first minimum wins */
    {
        int best = -1;
        for (int k = 0; k < 6; k++)
        {
            int cx = bmx + hex2[k + 1][0], cy = bmy + hex2[k + 1][1];
            int cost = me_sad_fpel(s, cx, cy) + me_mvcost(s, cx * 4, cy * 4);
            if (cost < bcost) { bcost = cost; best = k; }
        }
        if (best >= 0)
        {
            int dir = best;       /* index into hex2[dir+1] */
            bmx += hex2[dir + 1][0]; bmy += hex2[dir + 1][1];
            for (int i = (merange >> 1) - 1; i > 0 && bmx >= minx && bmx <= maxx && bmy >= miny && bmy <= maxy; i--)
            {
                int step = -1;
                for (int k = 0; k < 3; k++)
                {
                    int cx = bmx + hex2[dir + k][0], cy = bmy + hex2[dir + k][1];
                    int cost = me_sad_fpel(s, cx, cy) + me_mvcost(s, cx * 4, cy * 4);
                    if (cost < bcost) { bcost = cost; step = k; }
                }
                if (step < 0)
                    break;
                dir = mod6m1[dir + step - 1 + 1];
                bmx += hex2[dir + 1][0]; bmy += hex2[dir + 1][1];
            }
        }
    }
    /* square refine: 8 neighbours in the order of square1[1..8] */
    {
        int best = 0, ox = bmx, oy = bmy;
        for (int k = 1; k <= 8; k++)
        {
            int cx = ox + square1[k][0], cy = oy + square1[k][1];
            int cost = me_sad_fpel(s, cx, cy) + me_mvcost(s, cx * 4, cy * 4);
            if (cost < bcost) { bcost = cost; best = k; }
        }
        bmx += square1[best][0]; bmy += square1[best][1];
    }

    if (bprecost < bcost) { bmx = pmx; bmy = pmy; bcost = bprecost; }
    else { bmx *= 4; bmy *= 4; }

    if (!bcost)
        bcost = me_mvcost(s, bmx, bmy);
    else
    {
        int bdir = 0;
        for (int i = 1; i <= 4; i++)
        {
            int qx = bmx + square1[i][0] * 2, qy = bmy + square1[i][1] * 2;
            int cost = me_qpel_cost(s, qx, qy, 0) + me_mvcost(s, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmx += square1[bdir][0] * 2; bmy += square1[bdir][1] * 2;
        bcost = me_qpel_cost(s, bmx, bmy, 1) + me_mvcost(s, bmx, bmy);
        bdir = 0;
        for (int i = 1; i <= 4; i++)
        {
            int qx = bmx + square1[i][0], qy = bmy + square1[i][1];
            int cost = me_qpel_cost(s, qx, qy, 1) + me_mvcost(s, qx, qy);
            if (cost < bcost) { bcost = cost; bdir = i; }
        }
        bmx += square1[bdir][0]; bmy += square1[bdir][1];
    }
    *outx = bmx; *outy = bmy;
    return bcost;
}

/* ------------------------------------------------------------------------------------------
 * Weighted prediction analysis: LookaheadTLD::weightCostLuma / weightsAnalyse,
 * encoder/slicetype.cpp:338-488; WeightParam::setFromWeightAndOffset, common/slice.h:292-305
 * ---------------------------------------------------------------------------------------- */
static void ensure_wbuffer(ola_ctx* c, const ola_frame* f)
{
    if (c->wbuffer[0] && c->wplaneSize == f->g.planeSize) return;
    free(c->wbuffer[0]);
    c->wplaneSize = f->g.planeSize;
    c->wbuffer[0] = (pixel*)malloc((size_t)(4 * c->wplaneSize) * sizeof(pixel));
    for (int i = 1; i < 4; i++) c->wbuffer[i] = c->wbuffer[0] + i * c->wplaneSize;
}

static void weight_plane(ola_ctx* c, const ola_frame* ref, const ola_weight* w, int plane)
{
    int offset = w->offset << (ORACLE_DEPTH - 8);
    int round = w->denom ? 1 << (w->denom - 1) : 0;
    int correction = 14 - ORACLE_DEPTH;
    ola_weight_pp(ref->buffer[plane], c->wbuffer[plane], ref->g.stride, ref->g.stride, ref->g.paddedLines,
                  w->scale, round << correction, w->denom + correction, offset);
}

void ola_apply_weight(ola_ctx* c, ola_frame* ref, const ola_weight* w)
{
    ensure_wbuffer(c, ref);
    for (int i = 0; i < 4; i++) weight_plane(c, ref, w, i);
}

uint32_t ola_weight_cost_luma(ola_ctx* c, ola_frame* fenc, ola_frame* ref, const ola_weight* w)
{
    const pixel* src = ref->plane[0];
    intptr_t stride = fenc->g.stride;
    if (w && w->present)
    {
        ensure_wbuffer(c, ref);
        weight_plane(c, ref, w, 0);
        src = c->wbuffer[0] + fenc->g.padOffset;
    }
    uint32_t cost = 0;
    int mb = 0;
    for (int y = 0; y < fenc->g.lines; y += 8)
        for (int x = 0; x < fenc->g.width; x += 8, mb++)
        {
            intptr_t off = (intptr_t)y * stride + x;
            int satd = ola_satd8x8(src + off, stride, fenc->plane[0] + off, stride);
            cost += imin(satd, fenc->intraCost[mb]);
        }
    return cost;
}

void ola_weights_analyse(ola_ctx* c, ola_frame* fenc, ola_frame* ref, ola_weight* out)
{
    static const float epsilon = 1.f / 128.f;
    int deltaIndex = fenc->frameNum - ref->frameNum;
    out->present = 0; out->scale = 0; out->denom = 0; out->offset = 0;
    ensure_wbuffer(c, fenc);

    float guessScale, fencMean, refMean;
    if (fenc->wp_ssd[0] && ref->wp_ssd[0])
        guessScale = sqrtf((float)fenc->wp_ssd[0] / ref->wp_ssd[0]);
    else
        guessScale = 1.0f;
    fencMean = (float)fenc->wp_sum[0] / (fenc->g.lines * fenc->g.width) / (1 << (ORACLE_DEPTH - 8));
    refMean = (float)ref->wp_sum[0] / (fenc->g.lines * fenc->g.width) / (1 << (ORACLE_DEPTH - 8));

    if (fabsf(refMean - fencMean) < 0.5f && fabsf(1.f - guessScale) < epsilon)
        return;

    int minoff = 0, minscale, mindenom;
    unsigned int minscore = 0, origscore = 1;
    int found = 0;

    /* setFromWeightAndOffset((int)(guessScale * 128 + 0.5f), 0, 7, true) */
    ola_weight wp;
    wp.present = 0;
    wp.offset = 0;
    wp.denom = 7;
    wp.scale = (int)(guessScale * 128 + 0.5f);
    while (wp.denom > 0 && wp.scale > 127) { wp.denom--; wp.scale >>= 1; }
    wp.scale = imin(wp.scale, 127);
    mindenom = wp.denom;
    minscale = wp.scale;

    origscore = minscore = ola_weight_cost_luma(c, fenc, ref, &wp); /* bPresentFlag still false: unweighted */
    if (!minscore)
        return;

    unsigned int s = 0;
    int curScale = minscale;
    int curOffset = (int)(fencMean - refMean * curScale / (1 << mindenom) + 0.5f);
    if (curOffset < -128 || curOffset > 127)
    {
        curOffset = iclip(-128, 127, curOffset);
        curScale = (int)((1 << mindenom) * (fencMean - curOffset) / refMean + 0.5f);
        curScale = iclip(0, 127, curScale);
    }
    wp.present = 1; wp.scale = curScale; wp.denom = mindenom; wp.offset = curOffset;
    s = ola_weight_cost_luma(c, fenc, ref, &wp);
    if (s < minscore) { minscore = s; minscale = curScale; minoff = curOffset; found = 1; }

    while (mindenom > 0 && !(minscale & 1)) { mindenom--; minscale >>= 1; }

    if (!found || (minscale == 1 << mindenom && minoff == 0) || (float)minscore / origscore > 0.998f)
        return;
    out->present = 1; out->scale = minscale; out->denom = mindenom; out->offset = minoff;
    fenc->weightedCostDelta[deltaIndex] = minscore / origscore; /* unsigned integer division, as in the reference */
    ola_apply_weight(c, ref, out);
}

/* ------------------------------------------------------------------------------------------
 * CostEstimateGroup::estimateCUCost, encoder/slicetype.cpp:2068-2225
 * ---------------------------------------------------------------------------------------- */
typedef struct slice_acc { int64_t costEst, costEstAq; int intraMbs; } slice_acc;

static void estimate_cu(ola_ctx* c, ola_frame* fenc, ola_frame* fref0, ola_frame* fref1, const ref_planes* wref0,
                        int cuX, int cuY, int d0, int d1, const int doSearch[2], int lastRow, slice_acc* acc)
{
    const ola_geom* g = &fenc->g;
    const int W = g->wCU, H = g->hCU;
    const int bBidir = d1 > 0;
    const int cuXY = cuX + cuY * W;
    const intptr_t pelOffset = CU * cuX + (intptr_t)CU * cuY * g->stride;
    const int listDist[2] = { d0 - 1, d1 - 1 };
    pixel fencBlk[64];
    for (int y = 0; y < 8; y++)
        memcpy(fencBlk + 8 * y, fenc->plane[0] + pelOffset + (intptr_t)y * g->stride, 8 * sizeof(pixel));

    ref_planes r0 = { { fref0->plane[0], fref0->plane[1], fref0->plane[2], fref0->plane[3] }, g->stride };
    ref_planes r1 = { { fref1->plane[0], fref1->plane[1], fref1->plane[2], fref1->plane[3] }, g->stride };

    int bcost = OLA_COST_MAX, listused = 0;
    int minx = -cuX * CU - 8, miny = -cuY * CU - 8;
    int maxx = (W - cuX - 1) * CU + 8, maxy = (H - cuY - 1) * CU + 8;

    for (int i = 0; i < 1 + bBidir; i++)
    {
        int32_t* fencCost = &fenc->mvCosts[i][listDist[i]][cuXY];
        if (!doSearch[i])
        {
            if (*fencCost < bcost) { bcost = *fencCost; listused = i + 1; }
            continue;
        }
        ola_mv* fencMV = &fenc->mvs[i][listDist[i]][cuXY];
        const ref_planes* fref = i ? &r1 : wref0;
        int skipCost = 0x7fffffff;

        ola_mv mvc[4];
        int numc = 0;
        if (cuX < W - 1) mvc[numc++] = fencMV[1];
        if (!lastRow)
        {
            mvc[numc++] = fencMV[W];
            if (cuX > 0) mvc[numc++] = fencMV[W - 1];
            if (cuX < W - 1) mvc[numc++] = fencMV[W + 1];
        }
        int mvpx = 0, mvpy = 0;
        if (numc)
        {
            int mvpcost = OLA_COST_MAX;
            for (int k = 0; k < numc; k++)
            {
                pixel blk[64];
                lowres_mc(fref, pelOffset, mvc[k].x, mvc[k].y, blk);
                c->nSatd++;
                int cost = ola_satd8x8(fencBlk, 8, blk, 8);
                if (cost < mvpcost) { mvpcost = cost; mvpx = mvc[k].x; mvpy = mvc[k].y; }
                if (!(mvpx | mvpy) && bBidir)
                    skipCost = cost;
            }
        }
        me_state s = { c, fencBlk, fref, pelOffset, mvpx, mvpy };
        int ox, oy;
        int cost = motion_estimate(&s, minx, miny, maxx, maxy, &ox, &oy);
        if (skipCost < 64 && skipCost < cost && bBidir)
        {
            cost = skipCost;
            ox = oy = 0;
        }
        *fencCost = cost;
        fencMV->x = (int16_t)ox; fencMV->y = (int16_t)oy;
        if (cost < bcost) { bcost = cost; listused = i + 1; }
    }

    if (bBidir)
    {
        pixel b0[64], b1[64], avg[64];
        ola_mv m0 = fenc->mvs[0][listDist[0]][cuXY], m1 = fenc->mvs[1][listDist[1]][cuXY];
        lowres_mc(&r0, pelOffset, m0.x, m0.y, b0);    /* un-weighted L0 reference on purpose */
        lowres_mc(&r1, pelOffset, m1.x, m1.y, b1);
        ola_pixelavg8x8(avg, 8, b0, 8, b1, 8);
        c->nSatd++;
        int bicost = ola_satd8x8(fencBlk, 8, avg, 8);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        ola_pixelavg8x8(avg, 8, r0.plane[0] + pelOffset, g->stride, r1.plane[0] + pelOffset, g->stride);
        c->nSatd++;
        bicost = ola_satd8x8(fencBlk, 8, avg, 8);
        if (bicost < bcost) { bcost = bicost; listused = 3; }
        bcost += 4;
    }
    else
    {
        bcost += 4;
        if (fenc->intraCost[cuXY] < bcost) { bcost = fenc->intraCost[cuXY]; listused = 0; }
    }

    int scored = (cuX > 0 && cuX < W - 1 && cuY > 0 && cuY < H - 1) || W <= 2 || H <= 2;
    int bcostAq = (scored && fenc->invQscale) ? ((bcost * fenc->invQscale[cuXY] + 128) >> 8) : bcost;
    if (scored)
    {
        acc->costEst += bcost;
        acc->costEstAq += bcostAq;
        if (!listused && !bBidir) acc->intraMbs++;
    }
    fenc->rowSatds[d0][d1][cuY] += bcostAq;
    fenc->lowresCosts[d0][d1][cuXY] = (uint16_t)(imin(bcost, LOWRES_COST_MASK) | (listused << LOWRES_COST_SHIFT));
}

/* ------------------------------------------------------------------------------------------
 * CostEstimateGroup::estimateFrameCost (non-cached branch) + processTasks coop slices,
 * encoder/slicetype.cpp:1950-1970,1977-2066
 * ---------------------------------------------------------------------------------------- */
int64_t ola_estimate(ola_ctx* c, ola_frame* fenc, ola_frame* ref0, ola_frame* ref1, int d0, int d1,
                     int search0, int search1, int sliced, int weightp, const ola_weight* weight, ola_weight* usedWeight)
{
    const ola_geom* g = &fenc->g;
    int doSearch[2];
    doSearch[0] = search0 >= 0 ? search0 : (d0 > 0 && fenc->mvs[0][d0 - 1][0].x == OLA_MV_SENTINEL);
    doSearch[1] = search1 >= 0 ? search1 : (d1 > 0 && fenc->mvs[1][d1 - 1][0].x == OLA_MV_SENTINEL);
    c->nSad = c->nSatd = 0;

    ola_weight w = { 0, 0, 0, 0 };
    if (weight)
    {
        w = *weight;
        if (w.present) ola_apply_weight(c, ref0, &w);
    }
    else if (weightp && doSearch[0])
        ola_weights_analyse(c, fenc, ref0, &w);
    if (usedWeight) *usedWeight = w;

    ref_planes wref0;
    wref0.stride = g->stride;
    for (int i = 0; i < 4; i++)
        wref0.plane[i] = w.present ? c->wbuffer[i] + g->padOffset : ref0->plane[i];

    slice_acc total = { 0, 0, 0 };
    int useSlices = sliced && c->numCoopSlices > 1 && (d1 > 0 || doSearch[0] || doSearch[1]);
    int nSlices = useSlices ? c->numCoopSlices : 1;
    for (int sl = 0; sl < nSlices; sl++)
    {
        int firstY = useSlices ? c->numRowsPerSlice * sl : 0;
        int lastY = (!useSlices || sl == nSlices - 1) ? g->hCU - 1 : c->numRowsPerSlice * (sl + 1) - 1;
        int lastRow = 1;
        for (int cuY = lastY; cuY >= firstY; cuY--)
        {
            fenc->rowSatds[d0][d1][cuY] = 0;
            for (int cuX = g->wCU - 1; cuX >= 0; cuX--)
                estimate_cu(c, fenc, ref0, ref1, &wref0, cuX, cuY, d0, d1, doSearch, lastRow, &total);
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


/* ================================================================================================
 * cuTree propagation
 * ============================================================================================== */

/* estimateCUPropagateCost, common/pixel.cpp:848-874 (the #else branch, the one compiled).  The order of the
 * double operations is the one the reference's object code performs (checked against its disassembly: the
 * int32 product intraCost*invQscale wraps, then ((double)product * (fpsFactor/256) + propagateIn) *
 * (intra - min(intra, inter)) / intra + 0.5, truncated; every step is an IEEE-754 round-to-nearest operation,
 * no contraction), so plain C without -ffast-math reproduces it. */
void ola_propagate_cost(int* dst, const uint16_t* propagateIn, const int32_t* intraCosts, const uint16_t* interCosts,
                        const int32_t* invQscales, const double* fpsFactor, int len)
{
    volatile double fps = *fpsFactor * (1.0 / 256);
    for (int i = 0; i < len; i++)
    {
        int intra = intraCosts[i];
        int inter = interCosts[i] & 0x3FFF;                       /* LOWRES_COST_MASK, slicetype.h:41 */
        if (inter > intra) inter = intra;
        int32_t prod = (int32_t)((uint32_t)intra * (uint32_t)invQscales[i]);
        volatile double amount = (double)prod * fps;
        amount = amount + (double)propagateIn[i];
        volatile double r = amount * (double)(intra - inter);
        r = r / (double)intra;
        r = r + 0.5;
        /* cvttsd2si: out-of-range and NaN give INT_MIN */
        dst[i] = (r >= -2147483648.0 && r < 2147483648.0) ? (int)r : (int)0x80000000;
    }
}

/* memset(frames[x]->propagateCost, 0, ...) in Lookahead::cuTree, slicetype.cpp:1668-1701 */
void ola_cutree_zero(ola_frame* f)
{
    memset(f->propagateCost, 0, (size_t)f->g.nCU * sizeof(uint16_t));
}

static double ola_clip_duration(double f)     /* CLIP_DURATION, slicetype.cpp:44 */
{
    return f < 0.01 ? 0.01 : (f > 1.00 ? 1.00 : f);
}

#define OLA_CLIP_ADD(s, x) (s) = (uint16_t)((int)(s) + (x) < 65535 ? (int)(s) + (x) : 65535)

/* Lookahead::estimateCUPropagate, slicetype.cpp:1741-1839 (without the VBV-only cuTreeFinish at its end, which
 * the caller issues).  fenc = frames[b], ref0 = frames[p0], ref1 = frames[p1]. */
void ola_estimate_cu_propagate(ola_frame* fenc, ola_frame* ref0, ola_frame* ref1, int d0, int d1, int referenced,
                               double averageDuration, int fpsNum, int fpsDenom, int weightedBiPred)
{
    const int wCU = fenc->g.wCU, hCU = fenc->g.hCU;
    uint16_t* refCosts[2] = { ref0->propagateCost, ref1->propagateCost };
    int32_t distScaleFactor = ((d0 << 8) + ((d0 + d1) >> 1)) / (d0 + d1);
    int32_t bipredWeight = weightedBiPred ? 64 - (distScaleFactor >> 2) : 32;
    int32_t bipredWeights[2] = { bipredWeight, 64 - bipredWeight };
    int listDist[2] = { d0 - 1, d1 - 1 };
    int* scratch = (int*)calloc((size_t)wCU, sizeof(int));
    uint16_t* propagateCost = fenc->propagateCost;
    double fpsFactor = ola_clip_duration((double)fpsDenom / fpsNum) / ola_clip_duration(averageDuration);
    const uint16_t* lowresCosts = fenc->lowresCosts[d0][d1];

    if (!referenced)
        memset(fenc->propagateCost, 0, (size_t)wCU * sizeof(uint16_t));

    for (int blocky = 0; blocky < hCU; blocky++)
    {
        int cuIndex = blocky * wCU;
        ola_propagate_cost(scratch, propagateCost, fenc->intraCost + cuIndex, lowresCosts + cuIndex,
                           fenc->invQscale + cuIndex, &fpsFactor, wCU);
        if (referenced)
            propagateCost += wCU;
        for (int blockx = 0; blockx < wCU; blockx++, cuIndex++)
        {
            int32_t amount = scratch[blockx];
            if (amount <= 0) continue;                             /* intra block */
            int32_t listsUsed = lowresCosts[cuIndex] >> 14;
            for (int list = 0; list < 2; list++)
            {
                if (!((listsUsed >> list) & 1)) continue;
                int32_t listamount = amount;
                if (listsUsed == 3)
                    listamount = (listamount * bipredWeights[list] + 32) >> 6;
                const ola_mv* mvs = fenc->mvs[list][listDist[list]];
                if (!mvs[cuIndex].x && !mvs[cuIndex].y)
                {
                    OLA_CLIP_ADD(refCosts[list][cuIndex], listamount);
                    continue;
                }
                int32_t x = mvs[cuIndex].x, y = mvs[cuIndex].y;
                int32_t cux = (x >> 5) + blockx, cuy = (y >> 5) + blocky;
                int32_t idx0 = cux + cuy * wCU;
                x &= 31; y &= 31;
                int32_t wgt[4] = { (32 - y) * (32 - x), (32 - y) * x, y * (32 - x), y * x };
                for (int k = 0; k < 4; k++)
                {
                    int tx = cux + (k & 1), ty = cuy + (k >> 1);
                    /* both the all-inside branch and the per-corner branch reduce to: the corner is a CU of the frame */
                    if (tx >= 0 && tx < wCU && ty >= 0 && ty < hCU)
                        OLA_CLIP_ADD(refCosts[list][idx0 + (k & 1) + (k >> 1) * wCU], (listamount * wgt[k] + 512) >> 10);
                }
            }
        }
    }
    free(scratch);
}

/* Lookahead::cuTreeFinish, slicetype.cpp:1844-1862.  cuTreeStrength = 5.0 * (1.0 - qCompress) (slicetype.cpp:514). */
void ola_cutree_finish(ola_frame* f, double averageDuration, int fpsNum, int fpsDenom, int ref0Distance, double cuTreeStrength)
{
    int fpsFactor = (int)(ola_clip_duration(averageDuration) / ola_clip_duration((double)fpsDenom / fpsNum) * 256);
    double weightdelta = 0.0;
    if (ref0Distance && f->weightedCostDelta[ref0Distance - 1] > 0)
        weightdelta = 1.0 - f->weightedCostDelta[ref0Distance - 1];
    for (int i = 0; i < f->g.nCU; i++)
    {
        int intracost = (f->intraCost[i] * f->invQscale[i] + 128) >> 8;
        if (intracost)
        {
            int propagateCost = (f->propagateCost[i] * fpsFactor + 128) >> 8;
            double log2_ratio = log2((double)(intracost + propagateCost)) - log2((double)intracost) + weightdelta;
            f->qpCuTreeOffset[i] = f->qpAqOffset[i] - cuTreeStrength * log2_ratio;
        }
    }
}


/* ================================================================================================
 * full-resolution PU primitives
 * ============================================================================================== */

/* sad<lx, ly>, common/pixel.cpp:39-55 */
int ola_pu_sad(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int sum = 0;
    for (int y = 0; y < h; y++, a += sa, b += sb)
        for (int x = 0; x < w; x++)
            sum += abs((int)a[x] - (int)b[x]);
    return sum;
}

/* pu[LUMA_WxH].satd, common/pixel.cpp:979-1003: satd8<w, h> (sum of satd_8x4 over 8x4 tiles, :232-242) for widths that are
 * multiples of 8, satd4<w, h> (sum of satd_4x4 over 4x4 tiles, :220-230) for widths 4 and 12.  satd_8x4 (:192-218) halves the
 * packed sum of its two 4x4 Hadamard abs-sums; each of those is even (the coefficients of a 4x4 Hadamard sum to 16 * d00), so
 * every shape is the sum over its 4x4 tiles of (abs-sum >> 1) = ola_satd4x4. */
int ola_pu_satd(int w, int h, const pixel* a, intptr_t sa, const pixel* b, intptr_t sb)
{
    int sum = 0;
    for (int y = 0; y < h; y += 4)
        for (int x = 0; x < w; x += 4)
            sum += ola_satd4x4(a + y * sa + x, sa, b + y * sb + x, sb);
    return sum;
}

/* ================================================================================================
 * explicit weighted-prediction analysis, pixel loops: encoder/weightPrediction.cpp
 * ============================================================================================== */

/* mcLuma, weightPrediction.cpp:59-90: a motion-compensated copy of the lowres reference, CU by CU, each vector clipped to the
 * picture + 8 samples.  planes[4] = lowresPlane[0..3] of the reference, mvs = cuCount (x, y) pairs. */
void ola_wp_mc_luma(pixel* const planes[4], intptr_t stride, int width, int lines, const int16_t* mvs, pixel* mcout)
{
    ref_planes r = { { planes[0], planes[1], planes[2], planes[3] }, stride };
    int cu = 0;
    for (int y = 0; y < lines; y += 8)
    {
        int16_t miny = (int16_t)((-y - 8) * 4), maxy = (int16_t)((lines - y - 1 + 8) * 4);
        for (int x = 0; x < width; x += 8, cu++)
        {
            int16_t minx = (int16_t)((-x - 8) * 4), maxx = (int16_t)((width - x - 1 + 8) * 4);
            int mx = mvs[2 * cu], my = mvs[2 * cu + 1];
            mx = mx < minx ? minx : (mx > maxx ? maxx : mx);
            my = my < miny ? miny : (my > maxy ? maxy : my);
            pixel blk[64];
            lowres_mc(&r, (intptr_t)y * stride + x, mx, my, blk);
            for (int i = 0; i < 8; i++)
                memcpy(mcout + (intptr_t)(y + i) * stride + x, blk + 8 * i, 8 * sizeof(pixel));
        }
    }
}

/* g_chromaFilter, common/constants.cpp:247-257 */
static const int16_t wp_chroma_filter[8][4] = {
    { 0, 64, 0, 0 }, { -2, 58, 10, -2 }, { -4, 54, 16, -2 }, { -6, 46, 28, -4 },
    { -4, 36, 36, -4 }, { -4, 28, 46, -6 }, { -2, 16, 54, -4 }, { -2, 10, 58, -2 } };

/* mcChroma, weightPrediction.cpp:92-166, for 4:2:0 (8x8 chroma blocks; the four interpolation cases are
 * interp_horiz_pp_c / interp_vert_pp_c / interp_horiz_ps_c + interp_vert_sp_c with N = 4, common/ipfilter.cpp:80-284).
 * src points at sample (0,0) of a border-extended chroma plane.  The availability test and the vector index use the SAMPLE
 * position against the lowres CU counts, as the reference does (:113-121). */
void ola_wp_mc_chroma(const pixel* src, intptr_t stride, const int16_t* mvs, int lowresWidthInCU, int lowresHeightInCU,
                      int height, int width, pixel* mcout)
{
    const int maxVal = (1 << ORACLE_DEPTH) - 1;
    const int headRoom = 14 - ORACLE_DEPTH;
    for (int y = 0; y < height; y += 8)
    {
        int cu = y * lowresWidthInCU;
        int16_t miny = (int16_t)((-y - 8) * 4), maxy = (int16_t)((height - y - 1 + 8) * 4);
        for (int x = 0; x < width; x += 8, cu++)
        {
            intptr_t pixoff = (intptr_t)y * stride + x;
            if (x < lowresWidthInCU && y < lowresHeightInCU)
            {
                int16_t mx = mvs[2 * cu], my = mvs[2 * cu + 1];
                mx = (int16_t)(mx << 1); my = (int16_t)(my << 1);       /* lowres -> full resolution */
                mx >>= 1; my >>= 1;                                       /* -> 4:2:0 chroma */
                int16_t minx = (int16_t)((-x - 8) * 4), maxx = (int16_t)((width - x - 1 + 8) * 4);
                mx = mx < minx ? minx : (mx > maxx ? maxx : mx);
                my = my < miny ? miny : (my > maxy ? maxy : my);
                const pixel* temp = src + pixoff + (intptr_t)(my >> 2) * stride + (mx >> 2);
                int xFrac = mx & 7, yFrac = my & 7;
                const int16_t* cx = wp_chroma_filter[xFrac];
                const int16_t* cy = wp_chroma_filter[yFrac];
                for (int r = 0; r < 8; r++)
                    for (int c = 0; c < 8; c++)
                    {
                        const pixel* p = temp + (intptr_t)r * stride + c;
                        int val;
                        if (!(xFrac | yFrac))
                            val = p[0];
                        else if (!yFrac)
                        {
                            int sum = p[-1] * cx[0] + p[0] * cx[1] + p[1] * cx[2] + p[2] * cx[3];
                            int16_t v = (int16_t)((sum + 32) >> 6);
                            val = v < 0 ? 0 : (v > maxVal ? maxVal : v);
                        }
                        else if (!xFrac)
                        {
                            int sum = p[-stride] * cy[0] + p[0] * cy[1] + p[stride] * cy[2] + p[2 * stride] * cy[3];
                            int16_t v = (int16_t)((sum + 32) >> 6);
                            val = v < 0 ? 0 : (v > maxVal ? maxVal : v);
                        }
                        else
                        {
                            int shiftH = 6 - headRoom, offH = -(8192 << shiftH);
                            int shiftV = 6 + headRoom, offV = (1 << (shiftV - 1)) + (8192 << 6);
                            int sum = 0;
                            for (int k = 0; k < 4; k++)
                            {
                                const pixel* q = p + (intptr_t)(k - 1) * stride;
                                int h = q[-1] * cx[0] + q[0] * cx[1] + q[1] * cx[2] + q[2] * cx[3];
                                int16_t imm = (int16_t)((h + offH) >> shiftH);
                                sum += imm * cy[k];
                            }
                            int16_t v = (int16_t)((sum + offV) >> shiftV);
                            val = v < 0 ? 0 : (v > maxVal ? maxVal : v);
                        }
                        mcout[pixoff + (intptr_t)r * stride + c] = (pixel)val;
                    }
            }
            else
                for (int r = 0; r < 8; r++)
                    memcpy(mcout + pixoff + (intptr_t)r * stride, src + pixoff + (intptr_t)r * stride, 8 * sizeof(pixel));
        }
    }
}

/* weightCost, weightPrediction.cpp:168-220 (luma, and chroma of 4:2:0 / 4:2:2): the reference plane, weighted when
 * weighted != 0 (weight_pp_c into weightTemp, which must hold stride * height samples), against the source in 8x8 SATDs;
 * intraCost != NULL (luma) limits each of them. */
uint32_t ola_wp_cost(const pixel* fenc, const pixel* ref, pixel* weightTemp, intptr_t stride, int width, int height,
                     const int32_t* intraCost, int weighted, int scale, int denom, int offset)
{
    if (weighted)
    {
        int off = offset << (ORACLE_DEPTH - 8);
        int round = denom ? 1 << (denom - 1) : 0;
        int correction = 14 - ORACLE_DEPTH;
        int pwidth = ((width + 15) >> 4) << 4;
        ola_weight_pp(ref, weightTemp, stride, pwidth, height, scale, round << correction, denom + correction, off);
        ref = weightTemp;
    }
    uint32_t cost = 0;
    int cu = 0;
    for (int y = 0; y < height; y += 8)
        for (int x = 0; x < width; x += 8, cu++)
        {
            int cmp = ola_satd8x8(ref + (intptr_t)y * stride + x, stride, fenc + (intptr_t)y * stride + x, stride);
            cost += intraCost ? (uint32_t)imin(cmp, intraCost[cu]) : (uint32_t)cmp;
        }
    return cost;
}

/* ================================================================================================
 * full-resolution motion search of one PU: MotionEstimate::motionEstimate with
 * ref->isLowres == false, encoder/motion.cpp:571-1172, all five integer patterns (DIA :650-668, HEX :670-742,
 * UMH :744-927, STAR :929-1037 + StarPatternSearch :329-569, FULL :1039-1071), the predictor candidates (:627-641) and
 * the sub-pel refinement ladder (:1085-1168, workload[] :45-55) on subpelCompare (:1174-1203: luma only, the lookahead's
 * setSourcePU variant :165-181 has no chroma), whose fractional blocks come from the 8-tap filters
 * interp_horiz_pp_c / interp_vert_pp_c / interp_hv_pp_c (common/ipfilter.cpp:80-119,166-205,366-372).
 * Written as a list-driven machine ("measure these points in this order, keep the first strict minimum") instead of the
 * reference's macro ladders; decisions are the same, point for point.
 * ============================================================================================== */

/* g_lumaFilter, common/constants.cpp:239-245 */
static const int16_t pme_luma_filter[4][8] = {
    { 0, 0, 0, 64, 0, 0, 0, 0 }, { -1, 4, -10, 58, 17, -5, 1, 0 }, { -1, 4, -11, 40, 40, -11, 4, -1 }, { 0, 1, -5, 17, 58, -10, 4, -1 } };

typedef struct pme_state
{
    int w, h, sizeScale;                 /* sizeScale[partEnum] = (h * h) >> 4, motion.cpp:121-151 */
    const pixel* fenc; intptr_t fs;      /* the PU in the source plane */
    const pixel* fref; intptr_t rs;      /* fpelPlane[0] + blockOffset */
    const uint16_t* lut;                 /* centre of the mvcost table */
    int16_t mvpx, mvpy;                  /* setMVP(qmvp): the UNCLIPPED predictor */
    int16_t minx, miny, maxx, maxy;      /* full-pel search window */
    int16_t bx, by; int bcost;           /* best full-pel vector so far */
} pme_state;

static int pme_mvcost(const pme_state* s, int16_t qx, int16_t qy)
{
    return (uint16_t)(s->lut[qx - s->mvpx] + s->lut[qy - s->mvpy]);    /* BitCost::mvcost returns uint16_t, bitcost.h:45 */
}
static int pme_fpel_cost(const pme_state* s, int mx, int my)
{
    return ola_pu_sad(s->w, s->h, s->fenc, s->fs, s->fref + mx + (intptr_t)my * s->rs, s->rs) + pme_mvcost(s, (int16_t)(mx << 2), (int16_t)(my << 2));
}
static int pme_in_range(const pme_state* s, int x, int y) { return x >= s->minx && x <= s->maxx && y >= s->miny && y <= s->maxy; }
/* COST_MV (:225-231): returns 1 when the point became the best */
static int pme_try(pme_state* s, int mx, int my)
{
    int c = pme_fpel_cost(s, mx, my);
    if (c < s->bcost) { s->bcost = c; s->bx = (int16_t)mx; s->by = (int16_t)my; return 1; }
    return 0;
}

/* subpelCompare, :1174-1203, luma part */
static int pme_subpel(const pme_state* s, int16_t qx, int16_t qy, int useSatd)
{
    const pixel* r = s->fref + (qx >> 2) + (intptr_t)(qy >> 2) * s->rs;
    const int xf = qx & 3, yf = qy & 3, w = s->w, h = s->h;
    if (!(xf | yf))
        return useSatd ? ola_pu_satd(w, h, s->fenc, s->fs, r, s->rs) : ola_pu_sad(w, h, s->fenc, s->fs, r, s->rs);
    pixel buf[64 * 64];
    if (!yf || !xf)
    {
        /* one 8-tap pass, rounded to pixels: shift 6, offset 32 */
        const int16_t* c = pme_luma_filter[yf ? yf : xf];
        const intptr_t step = yf ? s->rs : 1;
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
            {
                const pixel* p = r + x + (intptr_t)y * s->rs - 3 * step;
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += p[k * step] * c[k];
                int16_t v = (int16_t)((sum + 32) >> 6);
                buf[y * 64 + x] = clip_pixel(v);
            }
    }
    else
    {
        /* horizontal pass into 14-bit intermediates over h + 7 rows, then the vertical pass */
        int16_t mid[(64 + 7) * 64];
        const int head = 14 - ORACLE_DEPTH, sh1 = 6 - head, off1 = -(8192 << sh1);
        const int sh2 = 6 + head, off2 = (1 << (sh2 - 1)) + (8192 << 6);
        const int16_t* cx = pme_luma_filter[xf];
        const int16_t* cy = pme_luma_filter[yf];
        for (int y = 0; y < h + 7; y++)
            for (int x = 0; x < w; x++)
            {
                const pixel* p = r + x - 3 + (intptr_t)(y - 3) * s->rs;
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += p[k] * cx[k];
                mid[y * w + x] = (int16_t)((sum + off1) >> sh1);
            }
        for (int y = 0; y < h; y++)
            for (int x = 0; x < w; x++)
            {
                int sum = 0;
                for (int k = 0; k < 8; k++) sum += mid[(y + k) * w + x] * cy[k];
                int16_t v = (int16_t)((sum + off2) >> sh2);
                buf[y * 64 + x] = clip_pixel(v);
            }
    }
    return useSatd ? ola_pu_satd(w, h, s->fenc, s->fs, buf, 64) : ola_pu_sad(w, h, s->fenc, s->fs, buf, 64);
}

static const int8_t pme_hex2[8][2] = { { -1, -2 }, { -2, 0 }, { -1, 2 }, { 1, 2 }, { 2, 0 }, { 1, -2 }, { -1, -2 }, { -2, 0 } };
static const int8_t pme_square1[9][2] = { { 0, 0 }, { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 }, { -1, -1 }, { -1, 1 }, { 1, -1 }, { 1, 1 } };
static const int8_t pme_hex4[16][2] = { { 0, -4 }, { 0, 4 }, { -2, -3 }, { 2, -3 }, { -4, -2 }, { 4, -2 }, { -4, -1 }, { 4, -1 },
                                       { -4, 0 }, { 4, 0 }, { -4, 1 }, { 4, 1 }, { -4, 2 }, { 4, 2 }, { -2, 3 }, { 2, 3 } };
/* offsets[] of the two-point search, :70-80 */
static const int8_t pme_two_point[16][2] = { { -1, 0 }, { 0, -1 }, { -1, -1 }, { 1, -1 }, { -1, 0 }, { 1, 0 }, { -1, 1 }, { -1, -1 },
                                            { 1, -1 }, { 1, 1 }, { -1, 0 }, { 0, 1 }, { -1, 1 }, { 1, 1 }, { 1, 0 }, { 0, 1 } };

/* n points around (ox, oy), measured in order; the first strict minimum below bcost wins (COST_MV_X4 :266-283) */
static void pme_try_around(pme_state* s, int ox, int oy, const int8_t (*d)[2], int n)
{
    for (int i = 0; i < n; i++) pme_try(s, ox + d[i][0], oy + d[i][1]);
}

/* the HEX pattern, :670-742: hexagon walk, then the 8-point square */
static void pme_hex(pme_state* s, int merange)
{
    int dir = -1, best = s->bcost;
    for (int k = 0; k < 6; k++)
    {
        /* order (-2,0) (-1,2) (1,2) (2,0) (1,-2) (-1,-2) = hex2[1..6] */
        int c = pme_fpel_cost(s, s->bx + pme_hex2[k + 1][0], s->by + pme_hex2[k + 1][1]);
        if (c < best) { best = c; dir = k; }
    }
    if (dir >= 0)
    {
        s->bcost = best;
        s->bx += pme_hex2[dir + 1][0]; s->by += pme_hex2[dir + 1][1];
        for (int i = (merange >> 1) - 1; i > 0 && pme_in_range(s, s->bx, s->by); i--)
        {
            int step = -1;
            for (int k = 0; k < 3; k++)
            {
                int c = pme_fpel_cost(s, s->bx + pme_hex2[dir + k][0], s->by + pme_hex2[dir + k][1]);
                if (c < best) { best = c; step = k; }
            }
            if (step < 0) break;
            s->bcost = best;
            dir += step - 1;                       /* dir += (bcost & 7) - 2 with tags 1..3 */
            dir = (dir + 6) % 6;                   /* mod6m1[dir + 1] */
            s->bx += pme_hex2[dir + 1][0]; s->by += pme_hex2[dir + 1][1];
        }
    }
    int sq = 0;
    for (int k = 1; k <= 8; k++)
    {
        int c = pme_fpel_cost(s, s->bx + pme_square1[k][0], s->by + pme_square1[k][1]);
        if (c < s->bcost) { s->bcost = c; sq = k; }
    }
    s->bx += pme_square1[sq][0]; s->by += pme_square1[sq][1];
}

/* CROSS (:303-327) around (ox, oy) */
static void pme_cross(pme_state* s, int ox, int oy, int start, int xmax, int ymax)
{
    int16_t i = (int16_t)start;
    if (xmax <= imin(s->maxx - ox, ox - s->minx))
        for (; i < xmax - 2; i += 4)
        {
            pme_try(s, ox + i, oy); pme_try(s, ox - i, oy); pme_try(s, ox + i + 2, oy); pme_try(s, ox - i - 2, oy);
        }
    for (; i < xmax; i += 2)
    {
        if (ox + i <= s->maxx) pme_try(s, ox + i, oy);
        if (ox - i >= s->minx) pme_try(s, ox - i, oy);
    }
    i = (int16_t)start;
    if (ymax <= imin(s->maxy - oy, oy - s->miny))
        for (; i < ymax - 2; i += 4)
        {
            pme_try(s, ox, oy + i); pme_try(s, ox, oy - i); pme_try(s, ox, oy + i + 2); pme_try(s, ox, oy - i - 2);
        }
    for (; i < ymax; i += 2)
    {
        if (oy + i <= s->maxy) pme_try(s, ox, oy + i);
        if (oy - i >= s->miny) pme_try(s, ox, oy - i);
    }
}

/* StarPatternSearch, :329-569.  Every ring is a list of points in one fixed order; a point is measured when the whole
 * ring lies inside the window, otherwise when ITS OWN one or two listed conditions hold (the reference tests only the
 * components that move: literal). */
static void pme_star_point(pme_state* s, int x, int y, int ringInside, int okA, int okB, int point, int dist, int* bPoint, int* bDist)
{
    if (!(ringInside || (okA && okB))) return;
    if (pme_try(s, x, y)) { *bPoint = point; *bDist = dist; }
}
static void pme_star_pattern(pme_state* s, int* bPoint, int* bDist, int earlyExitIters, int merange)
{
    const int ox = s->bx, oy = s->by;
    int rounds = 0;
    for (int dist = 1; dist <= 8; dist <<= 1)
    {
        const int16_t top = (int16_t)(oy - dist), bottom = (int16_t)(oy + dist), left = (int16_t)(ox - dist), right = (int16_t)(ox + dist);
        const int16_t top2 = (int16_t)(oy - (dist >> 1)), bottom2 = (int16_t)(oy + (dist >> 1)), left2 = (int16_t)(ox - (dist >> 1)), right2 = (int16_t)(ox + (dist >> 1));
        const int saved = s->bcost;
        const int in = top >= s->miny && left >= s->minx && right <= s->maxx && bottom <= s->maxy;
        pme_star_point(s, ox, top, in, top >= s->miny, 1, 2, dist, bPoint, bDist);
        if (dist > 1)
        {
            pme_star_point(s, left2, top2, in, top2 >= s->miny, left2 >= s->minx, 1, dist >> 1, bPoint, bDist);
            pme_star_point(s, right2, top2, in, top2 >= s->miny, right2 <= s->maxx, 3, dist >> 1, bPoint, bDist);
        }
        pme_star_point(s, left, oy, in, left >= s->minx, 1, 4, dist, bPoint, bDist);
        pme_star_point(s, right, oy, in, right <= s->maxx, 1, 5, dist, bPoint, bDist);
        if (dist > 1)
        {
            pme_star_point(s, left2, bottom2, in, bottom2 <= s->maxy, left2 >= s->minx, 6, dist >> 1, bPoint, bDist);
            pme_star_point(s, right2, bottom2, in, bottom2 <= s->maxy, right2 <= s->maxx, 8, dist >> 1, bPoint, bDist);
        }
        pme_star_point(s, ox, bottom, in, bottom <= s->maxy, 1, 7, dist, bPoint, bDist);
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
    for (int16_t dist = 16; dist <= (int16_t)merange; dist <<= 1)
    {
        const int16_t top = (int16_t)(oy - dist), bottom = (int16_t)(oy + dist), left = (int16_t)(ox - dist), right = (int16_t)(ox + dist);
        const int saved = s->bcost;
        const int in = top >= s->miny && left >= s->minx && right <= s->maxx && bottom <= s->maxy;
        pme_star_point(s, ox, top, in, top >= s->miny, 1, 0, dist, bPoint, bDist);
        pme_star_point(s, left, oy, in, left >= s->minx, 1, 0, dist, bPoint, bDist);
        pme_star_point(s, right, oy, in, right <= s->maxx, 1, 0, dist, bPoint, bDist);
        pme_star_point(s, ox, bottom, in, bottom <= s->maxy, 1, 0, dist, bPoint, bDist);
        for (int index = 1; index < 4; index++)
        {
            const int16_t yt = (int16_t)(top + (dist >> 2) * index), yb = (int16_t)(bottom - (dist >> 2) * index);
            const int16_t xl = (int16_t)(ox - (dist >> 2) * index), xr = (int16_t)(ox + (dist >> 2) * index);
            pme_star_point(s, xl, yt, in, yt >= s->miny, xl >= s->minx, 0, dist, bPoint, bDist);
            pme_star_point(s, xr, yt, in, yt >= s->miny, xr <= s->maxx, 0, dist, bPoint, bDist);
            pme_star_point(s, xl, yb, in, yb <= s->maxy, xl >= s->minx, 0, dist, bPoint, bDist);
            pme_star_point(s, xr, yb, in, yb <= s->maxy, xr <= s->maxx, 0, dist, bPoint, bDist);
        }
        if (s->bcost < saved) rounds = 0;
        else if (++rounds >= earlyExitIters) return;
    }
}
static void pme_two_points(pme_state* s, int point)
{
    /* bmv moves when the first point wins: the second is relative to the vector the pair started from */
    const int x0 = s->bx, y0 = s->by;
    for (int k = 0; k < 2; k++)
    {
        const int x = (int16_t)(x0 + pme_two_point[(point - 1) * 2 + k][0]), y = (int16_t)(y0 + pme_two_point[(point - 1) * 2 + k][1]);
        if (pme_in_range(s, x, y)) pme_try(s, x, y);
    }
}
/* the raster scan of STAR (:971-1003, step 5, with its `<< 3` in the fourth vector's mvcost) and FULL (:1039-1071, step 1) */
static void pme_raster(pme_state* s, int step, int fourthShift)
{
    for (int16_t ty = s->miny; ty <= s->maxy; ty += step)
        for (int16_t tx = s->minx; tx <= s->maxx; tx += step)
        {
            if (tx + step * 3 <= s->maxx)
            {
                for (int k = 0; k < 4; k++)
                {
                    if (k) tx += step;
                    const int sh = k == 3 ? fourthShift : 2;
                    int c = ola_pu_sad(s->w, s->h, s->fenc, s->fs, s->fref + tx + (intptr_t)ty * s->rs, s->rs) + pme_mvcost(s, (int16_t)(tx << sh), (int16_t)(ty << sh));
                    if (c < s->bcost) { s->bcost = c; s->bx = tx; s->by = ty; }
                }
            }
            else
                pme_try(s, tx, ty);
        }
}

/* UMH before its final hexagon walk, :744-927; returns 0 when the pattern ended early (no walk), 1 when the HEX walk
 * follows; *merange may come back widened (:789-833) */
static int pme_umh(pme_state* s, int16_t pmx, int16_t pmy, int16_t qmvpx, int16_t qmvpy, int numCandidates, const int16_t* mvc, int* merangeIO, int is64)
{
    static const int8_t dia1[4][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 } };
    static const int8_t octA[4][2] = { { 0, -2 }, { -1, -1 }, { 1, -1 }, { -2, 0 } }, octB[4][2] = { { 2, 0 }, { -1, 1 }, { 1, 1 }, { 0, 2 } };
    static const int8_t ringA[4][2] = { { -1, -2 }, { 1, -2 }, { -2, -1 }, { 2, -1 } }, ringB[4][2] = { { -2, 1 }, { 2, 1 }, { -1, 2 }, { 1, 2 } };
    static const int8_t corners[4][2] = { { -2, -2 }, { -2, 2 }, { 2, -2 }, { 2, 2 } };
    static const uint8_t range_mul[4][4] = { { 3, 3, 4, 4 }, { 3, 4, 4, 4 }, { 4, 4, 4, 5 }, { 4, 4, 5, 6 } };
    int merange = *merangeIO;
    int crossStart = 1;
    const int ucost1 = s->bcost;
    pme_try_around(s, pmx, pmy, dia1, 4);
    if (pmx | pmy) pme_try_around(s, 0, 0, dia1, 4);
    const int ucost2 = s->bcost;
    if ((s->bx | s->by) && !(s->bx == pmx && s->by == pmy)) pme_try_around(s, s->bx, s->by, dia1, 4);
    if (s->bcost == ucost2) crossStart = 3;
    int ox = s->bx, oy = s->by;
#define ME_THRESH(v) (s->bcost < (((v) >> 4) * s->sizeScale))
    if (s->bcost == ucost2 && ME_THRESH(2000))
    {
        pme_try_around(s, ox, oy, octA, 4);
        pme_try_around(s, ox, oy, octB, 4);
        if (s->bcost == ucost1 && ME_THRESH(500)) return 0;
        if (s->bcost == ucost2)
        {
            const int16_t range = (int16_t)(merange >> 1) | 1;
            pme_cross(s, ox, oy, 3, range, range);
            pme_try_around(s, ox, oy, ringA, 4);
            pme_try_around(s, ox, oy, ringB, 4);
            if (s->bcost == ucost2) return 0;
            crossStart = range + 2;
        }
    }
    if (numCandidates)
    {
        int mvd, denom = 1;
        if (numCandidates == 1)
            mvd = is64 ? 25 : abs(qmvpx - mvc[0]) + abs(qmvpy - mvc[1]);
        else
        {
            denom = numCandidates - 1;
            mvd = 0;
            if (!is64) { mvd = abs(qmvpx - mvc[0]) + abs(qmvpy - mvc[1]); denom++; }
            for (int i = 0; i < numCandidates - 1; i++)
                mvd += abs(mvc[2 * i] - mvc[2 * i + 2]) + abs(mvc[2 * i + 1] - mvc[2 * i + 3]);
        }
        const int sadCtx = ME_THRESH(1000) ? 0 : ME_THRESH(2000) ? 1 : ME_THRESH(4000) ? 2 : 3;
        const int mvdCtx = mvd < 10 * denom ? 0 : mvd < 20 * denom ? 1 : mvd < 40 * denom ? 2 : 3;
        merange = (merange * range_mul[mvdCtx][sadCtx]) >> 2;
    }
#undef ME_THRESH
    /* still centred where the small patterns were (the reference's own FIXME) */
    pme_cross(s, ox, oy, crossStart, merange, merange >> 1);
    pme_try_around(s, ox, oy, corners, 4);
    /* 16-point hexagon grid, scaled 1 .. merange / 4 */
    ox = s->bx; oy = s->by;
    uint16_t i = 1;
    do
    {
        if (4 * i > imin(imin(s->maxx - ox, ox - s->minx), imin(s->maxy - oy, oy - s->miny)))
        {
            for (int j = 0; j < 16; j++)
            {
                const int16_t x = (int16_t)(ox + (int16_t)(pme_hex4[j][0] * (int16_t)i)), y = (int16_t)(oy + (int16_t)(pme_hex4[j][1] * (int16_t)i));
                if (pme_in_range(s, x, y)) pme_try(s, x, y);
            }
        }
        else
        {
            int won = -1;
            for (int j = 0; j < 16; j++)
            {
                const int x = ox + pme_hex4[j][0] * i, y = oy + pme_hex4[j][1] * i;
                /* ADD_MVCOST reads the table rows directly (no uint16 wrap of the sum; the sum of two entries is < 65536) */
                int c = ola_pu_sad(s->w, s->h, s->fenc, s->fs, s->fref + x + (intptr_t)y * s->rs, s->rs) + s->lut[x * 4 - s->mvpx] + s->lut[y * 4 - s->mvpy];
                if (c < s->bcost) { s->bcost = c; won = j; }
            }
            if (won >= 0) { s->bx = (int16_t)(ox + i * pme_hex4[won][0]); s->by = (int16_t)(oy + i * pme_hex4[won][1]); }
        }
    }
    while (++i <= merange >> 2);
    *merangeIO = merange;
    return pme_in_range(s, s->bx, s->by);
}

/* searchMethod: 0 DIA, 1 HEX, 2 UMH, 3 STAR, 4 FULL (x265.h X265_*_SEARCH); subpelRefine 0..7.
 * fencPlane + offset = the PU in the source (stride fencStride); refPlane + offset = the co-located block of the
 * reference's fpelPlane[0] (stride refStride).  mvmin / mvmax in full-pel units, qmvp and mvc in quarter-pel units.
 * Returns the cost, writes the quarter-pel vector. */
int ola_motion_estimate_pu(int searchMethod, int subpelRefine, int w, int h, const pixel* fencPlane, intptr_t fencStride,
                           const pixel* refPlane, intptr_t refStride, intptr_t offset, const uint16_t* mvcostCentre,
                           const int16_t mvmin[2], const int16_t mvmax[2], const int16_t qmvp[2],
                           int numCandidates, const int16_t* mvc, int merange, int16_t outQMv[2])
{
    static const struct { int hpelIters, hpelDirs, qpelIters, qpelDirs, hpelSatd; } workload[8] = {
        { 1, 4, 0, 4, 0 }, { 1, 4, 1, 4, 0 }, { 1, 4, 1, 4, 1 }, { 2, 4, 1, 4, 1 }, { 2, 4, 2, 4, 1 }, { 1, 8, 1, 8, 1 }, { 2, 8, 1, 8, 1 }, { 2, 8, 2, 8, 1 } };
    pme_state S, *s = &S;
    s->w = w; s->h = h; s->sizeScale = (h * h) >> 4;
    s->fenc = fencPlane + offset; s->fs = fencStride;
    s->fref = refPlane + offset; s->rs = refStride;
    s->lut = mvcostCentre;
    s->mvpx = qmvp[0]; s->mvpy = qmvp[1];
    s->minx = mvmin[0]; s->miny = mvmin[1]; s->maxx = mvmax[0]; s->maxy = mvmax[1];
    const int16_t qminx = (int16_t)(mvmin[0] << 2), qminy = (int16_t)(mvmin[1] << 2), qmaxx = (int16_t)(mvmax[0] << 2), qmaxy = (int16_t)(mvmax[1] << 2);

    /* the clipped predictor: SAD without mvcost (:597-603) */
    int16_t pqx = qmvp[0] > qmaxx ? qmaxx : qmvp[0]; if (pqx < qminx) pqx = qminx;
    int16_t pqy = qmvp[1] > qmaxy ? qmaxy : qmvp[1]; if (pqy < qminy) pqy = qminy;
    int16_t bestPreX = pqx, bestPreY = pqy;
    int bprecost = pme_subpel(s, pqx, pqy, 0);
    const int16_t pmx = (int16_t)((pqx + 2) >> 2), pmy = (int16_t)((pqy + 2) >> 2);   /* pmv.roundToFPel() */
    s->bx = pmx; s->by = pmy;
    s->bcost = bprecost;
    if ((pqx | pqy) & 3)
        s->bcost = pme_fpel_cost(s, s->bx, s->by);
    if (pqx | pqy)
    {
        int c = pme_fpel_cost(s, 0, 0);
        if (c < s->bcost) { s->bcost = c; s->bx = s->by = 0; }
    }
    /* the other predictor candidates at quarter-pel precision (:627-641) */
    for (int i = 0; i < numCandidates; i++)
    {
        int16_t mx = mvc[2 * i] > qmaxx ? qmaxx : mvc[2 * i]; if (mx < qminx) mx = qminx;
        int16_t my = mvc[2 * i + 1] > qmaxy ? qmaxy : mvc[2 * i + 1]; if (my < qminy) my = qminy;
        if ((mx | my) && !(mx == pqx && my == pqy) && !(mx == bestPreX && my == bestPreY))
        {
            int c = pme_subpel(s, mx, my, 0) + pme_mvcost(s, mx, my);
            if (c < bprecost) { bprecost = c; bestPreX = mx; bestPreY = my; }
        }
    }

    int walk = 0;
    switch (searchMethod)
    {
    case 0:
    {
        /* diamond, radius 1, at most merange steps (:650-668); tags 1, 3, 4, 12 = up, down, left, right */
        static const int8_t dia[4][2] = { { 0, -1 }, { 0, 1 }, { -1, 0 }, { 1, 0 } };
        int i = merange;
        do
        {
            int step = -1, best = s->bcost;
            for (int k = 0; k < 4; k++)
            {
                int c = pme_fpel_cost(s, s->bx + dia[k][0], s->by + dia[k][1]);
                if (c < best) { best = c; step = k; }
            }
            if (step < 0) break;
            s->bcost = best;
            s->bx += dia[step][0]; s->by += dia[step][1];
        }
        while (--i && pme_in_range(s, s->bx, s->by));
        break;
    }
    case 1:
        walk = 1;
        break;
    case 2:
        walk = pme_umh(s, pmx, pmy, qmvp[0], qmvp[1], numCandidates, mvc, &merange, w == 64 && h == 64);
        break;
    case 3:
    {
        int bPoint = 0, bDist = 0;
        pme_star_pattern(s, &bPoint, &bDist, 3, merange);
        if (bDist == 1)
        {
            /* best distance 1: the two outer points next to the winning direction; stop if neither helps */
            if (!bPoint) break;
            const int saved = s->bcost;
            pme_two_points(s, bPoint);
            if (s->bcost == saved) break;
        }
        if (bDist > 5)
            pme_raster(s, 5, 3);
        while (bDist > 0)
        {
            bDist = 0; bPoint = 0;
            pme_star_pattern(s, &bPoint, &bDist, 32, merange);
            if (bDist == 1)
            {
                if (bPoint) pme_two_points(s, bPoint);
                break;
            }
        }
        break;
    }
    case 4:
        pme_raster(s, 1, 2);
        break;
    default:
        return -1;
    }
    if (walk) pme_hex(s, merange);

    int16_t qx, qy; int bcost;
    if (bprecost < s->bcost) { qx = bestPreX; qy = bestPreY; bcost = bprecost; }
    else { qx = (int16_t)(s->bx << 2); qy = (int16_t)(s->by << 2); bcost = s->bcost; }

    if (!bcost)
        bcost = pme_mvcost(s, qx, qy);                      /* zero residual: no sub-pel, but the vector's bits count (:1085-1090) */
    else
    {
        const int hs = workload[subpelRefine].hpelSatd;
        if (hs) bcost = pme_subpel(s, qx, qy, 1) + pme_mvcost(s, qx, qy);
        for (int pass = 0; pass < 2; pass++)
        {
            /* pass 0: half-pel steps with SAD or SATD; pass 1: quarter-pel steps with SATD */
            const int iters = pass ? workload[subpelRefine].qpelIters : workload[subpelRefine].hpelIters;
            const int dirs = pass ? workload[subpelRefine].qpelDirs : workload[subpelRefine].hpelDirs;
            const int mul = pass ? 1 : 2, satd = pass ? 1 : hs;
            if (pass && !hs) bcost = pme_subpel(s, qx, qy, 1) + pme_mvcost(s, qx, qy);
            for (int it = 0; it < iters; it++)
            {
                int bdir = 0;
                for (int i = 1; i <= dirs; i++)
                {
                    const int16_t cx = (int16_t)(qx + pme_square1[i][0] * mul), cy = (int16_t)(qy + pme_square1[i][1] * mul);
                    int c = pme_subpel(s, cx, cy, satd) + pme_mvcost(s, cx, cy);
                    if (c < bcost) { bcost = c; bdir = i; }
                }
                if (!bdir) break;
                qx = (int16_t)(qx + pme_square1[bdir][0] * mul); qy = (int16_t)(qy + pme_square1[bdir][1] * mul);
            }
        }
    }
    outQMv[0] = qx; outQMv[1] = qy;
    return bcost;
}

/* n searches of one PU shape: items as in include/x265cu.h's x265cu_me_item (same layout) */
void ola_motion_estimate_batch(int searchMethod, int subpelRefine, int w, int h, const pixel* fencPlane, intptr_t fencStride,
                               const pixel* refPlane, intptr_t refStride, const uint16_t* mvcostCentre, int n, const ola_me_item* items, ola_me_result* out)
{
    for (int i = 0; i < n; i++)
    {
        const ola_me_item* it = &items[i];
        out[i].cost = ola_motion_estimate_pu(searchMethod, subpelRefine, w, h, fencPlane, fencStride, refPlane, refStride, (intptr_t)it->offset, mvcostCentre,
                                             it->mvmin, it->mvmax, it->qmvp, it->numCandidates, &it->mvc[0][0], it->merange, out[i].mv);
    }
}
