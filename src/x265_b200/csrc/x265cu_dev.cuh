/* x265cu_dev.cuh -- device-side building blocks shared by the kernels of libx265cu.so (sm_100a).
 *
 * Lane mapping used everywhere a block of 8x8 samples is measured: a QUAD of 4 consecutive lanes
 * owns one 8x8 candidate block, lane (sub = lane & 3) owns its 4x4 sub-block at
 * (bx, by) = ((sub & 1) * 4, (sub >> 1) * 4).  SATD-8x8 in x265 is the sum of the four 4x4
 * Hadamard abs-sums (two satd_8x4 halves, common/pixel.cpp:192-242), so each lane transforms its
 * own 4x4 entirely in registers and a quad needs only two shuffles to finish; a warp measures up to
 * eight candidates per pass.  SAD uses the packed-byte/halfword video instructions
 * (__vsadu4/__vsadu2), rounded averages use __vavgu4/__vavgu2, which are exactly
 * pixelavg_pp's (a + b + 1) >> 1 (pixel.cpp:490-502).
 */
#ifndef X265CU_DEV_CUH
#define X265CU_DEV_CUH

#include <stdint.h>
#include <cuda_runtime.h>
#include "la_core.h"

#define FULL_MASK 0xffffffffu

/* ---- 4-sample rows, any alignment, 8- and 16-bit samples -------------------------------- */
template <typename P> struct Px;

template <> struct Px<uint8_t>
{
    struct Row4 { uint32_t v; };
    enum { PIXEL_MAX8 = 255 };
    static __device__ __forceinline__ Row4 load(const uint8_t* p)
    {
        uintptr_t a = (uintptr_t)p;
        const uint32_t* b = (const uint32_t*)(a & ~(uintptr_t)3);
        uint32_t sh = (uint32_t)(a & 3) * 8;
        Row4 r;
        r.v = __funnelshift_r(__ldg(b), __ldg(b + 1), sh);
        return r;
    }
    static __device__ __forceinline__ Row4 load_aligned(const uint8_t* p)
    {
        Row4 r; r.v = __ldg((const uint32_t*)p); return r;
    }
    /* 4 samples starting k (0..3) samples into the 8 samples u0:u1 */
    static __device__ __forceinline__ Row4 combine(Row4 u0, Row4 u1, int k) { Row4 r; r.v = __funnelshift_r(u0.v, u1.v, k * 8); return r; }
    static __device__ __forceinline__ Row4 avg(Row4 a, Row4 b) { Row4 r; r.v = __vavgu4(a.v, b.v); return r; }
    static __device__ __forceinline__ int sad(Row4 a, Row4 b) { return (int)__vsadu4(a.v, b.v); }
    static __device__ __forceinline__ void unpack(Row4 a, int v[4])
    {
        v[0] = a.v & 0xff; v[1] = (a.v >> 8) & 0xff; v[2] = (a.v >> 16) & 0xff; v[3] = a.v >> 24;
    }
    static __device__ __forceinline__ Row4 pack(const int v[4])
    {
        Row4 r; r.v = (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24); return r;
    }
};

template <> struct Px<uint16_t>
{
    struct Row4 { uint32_t lo, hi; };
    static __device__ __forceinline__ Row4 load(const uint16_t* p)
    {
        uintptr_t a = (uintptr_t)p;
        const uint32_t* b = (const uint32_t*)(a & ~(uintptr_t)3);
        uint32_t sh = (uint32_t)(a & 2) * 8;
        uint32_t w0 = __ldg(b), w1 = __ldg(b + 1), w2 = __ldg(b + 2);
        Row4 r;
        r.lo = __funnelshift_r(w0, w1, sh);
        r.hi = __funnelshift_r(w1, w2, sh);
        return r;
    }
    static __device__ __forceinline__ Row4 load_aligned(const uint16_t* p)
    {
        uint2 w = __ldg((const uint2*)p);
        Row4 r; r.lo = w.x; r.hi = w.y; return r;
    }
    static __device__ __forceinline__ Row4 combine(Row4 u0, Row4 u1, int k)
    {
        const uint32_t sh = (uint32_t)(k & 1) * 16;
        const uint32_t m0 = (k & 2) ? u0.hi : u0.lo, m1 = (k & 2) ? u1.lo : u0.hi, m2 = (k & 2) ? u1.hi : u1.lo;
        Row4 r; r.lo = __funnelshift_r(m0, m1, sh); r.hi = __funnelshift_r(m1, m2, sh); return r;
    }
    static __device__ __forceinline__ Row4 avg(Row4 a, Row4 b)
    {
        Row4 r; r.lo = __vavgu2(a.lo, b.lo); r.hi = __vavgu2(a.hi, b.hi); return r;
    }
    static __device__ __forceinline__ int sad(Row4 a, Row4 b) { return (int)(__vsadu2(a.lo, b.lo) + __vsadu2(a.hi, b.hi)); }
    static __device__ __forceinline__ void unpack(Row4 a, int v[4])
    {
        v[0] = a.lo & 0xffff; v[1] = a.lo >> 16; v[2] = a.hi & 0xffff; v[3] = a.hi >> 16;
    }
    static __device__ __forceinline__ Row4 pack(const int v[4])
    {
        Row4 r; r.lo = (uint32_t)v[0] | ((uint32_t)v[1] << 16); r.hi = (uint32_t)v[2] | ((uint32_t)v[3] << 16); return r;
    }
};

/* ---- the four hpel planes of one reference frame --------------------------------------- */
template <typename P> struct RefPlanes
{
    const P* p0;          /* sample (0,0) of plane 0; plane k at p0 + k * planeSize */
    int64_t planeSize;
    int stride;
};

/* 4x4 sub-block at sample position (x, y) of the lowres frame, displaced by the quarter-pel MV
 * (qx, qy): ReferencePlanes::lowresMC / lowresQPelCost, common/lowres.h:62-103 */
template <typename P>
__device__ __forceinline__ void mc_fetch4x4(const RefPlanes<P>& r, int x, int y, int qx, int qy, typename Px<P>::Row4 out[4])
{
    LaMcSrc m = la_mc_src(qx, qy);
    const P* a = r.p0 + (int64_t)m.planeA * r.planeSize + (int64_t)(y + m.ay) * r.stride + (x + m.ax);
    if (m.avg)
    {
        const P* b = r.p0 + (int64_t)m.planeB * r.planeSize + (int64_t)(y + m.by) * r.stride + (x + m.bx);
#pragma unroll
        for (int i = 0; i < 4; i++)
            out[i] = Px<P>::avg(Px<P>::load(a + (int64_t)i * r.stride), Px<P>::load(b + (int64_t)i * r.stride));
    }
    else
    {
#pragma unroll
        for (int i = 0; i < 4; i++)
            out[i] = Px<P>::load(a + (int64_t)i * r.stride);
    }
}

/* ---- per-lane 4x4 measures -------------------------------------------------------------- */
template <typename P>
__device__ __forceinline__ int sad4x4(const typename Px<P>::Row4 f[4], const typename Px<P>::Row4 r[4])
{
    return Px<P>::sad(f[0], r[0]) + Px<P>::sad(f[1], r[1]) + Px<P>::sad(f[2], r[2]) + Px<P>::sad(f[3], r[3]);
}

/* sum |H4 * D * H4| of the 4x4 difference block (not halved): hadamard of satd_8x4, pixel.cpp:143-242 */
__device__ __forceinline__ int hadamard4x4_abs(int d[4][4])
{
    int sum = 0;
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
        sum += abs(s01 + s23) + abs(d01 + d23) + abs(s01 - s23) + abs(d01 - d23);
    }
    return sum;
}

/* Row stage of the 4x4 Hadamard straight from packed samples with the integer dot-product
 * instructions: coefficient k of row y of H*(f - r) = dp(f_y, h_k) + dp(r_y, -h_k), h_k in {+1,-1}^4.
 * dp4a (8-bit samples) / dp2a (16-bit samples) are exact integer ops, so this is bit-equal to
 * unpack + subtract + butterflies while using a quarter of the instructions. */
__device__ __forceinline__ int dp4a_us(uint32_t a_u8x4, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u8x4), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_lo_us(uint32_t a_u16x2, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u16x2), "r"(b_s8x4), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi_us(uint32_t a_u16x2, uint32_t b_s8x4, int c)
{
    int d;
    asm("dp2a.hi.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a_u16x2), "r"(b_s8x4), "r"(c));
    return d;
}

/* +-1 patterns (byte i = sample i): h0 = ++++, h1 = +-+-, h2 = ++--, h3 = +--+ and their negations */
#define HAD_P0 0x01010101u
#define HAD_P1 0xFF01FF01u
#define HAD_P2 0xFFFF0101u
#define HAD_P3 0x01FFFF01u
#define HAD_N0 0xFFFFFFFFu
#define HAD_N1 0x01FF01FFu
#define HAD_N2 0x0101FFFFu
#define HAD_N3 0xFF0101FFu

__device__ __forceinline__ void had_row(Px<uint8_t>::Row4 f, Px<uint8_t>::Row4 r, int t[4])
{
    t[0] = dp4a_us(f.v, HAD_P0, dp4a_us(r.v, HAD_N0, 0));
    t[1] = dp4a_us(f.v, HAD_P1, dp4a_us(r.v, HAD_N1, 0));
    t[2] = dp4a_us(f.v, HAD_P2, dp4a_us(r.v, HAD_N2, 0));
    t[3] = dp4a_us(f.v, HAD_P3, dp4a_us(r.v, HAD_N3, 0));
}

__device__ __forceinline__ int dp2a4(Px<uint16_t>::Row4 a, uint32_t h, int c)
{
    return dp2a_hi_us(a.hi, h, dp2a_lo_us(a.lo, h, c));
}

__device__ __forceinline__ void had_row(Px<uint16_t>::Row4 f, Px<uint16_t>::Row4 r, int t[4])
{
    t[0] = dp2a4(f, HAD_P0, dp2a4(r, HAD_N0, 0));
    t[1] = dp2a4(f, HAD_P1, dp2a4(r, HAD_N1, 0));
    t[2] = dp2a4(f, HAD_P2, dp2a4(r, HAD_N2, 0));
    t[3] = dp2a4(f, HAD_P3, dp2a4(r, HAD_N3, 0));
}

/* sum |H4 * (F - R) * H4| of this lane's 4x4 block (not halved) */
template <typename P>
__device__ __forceinline__ int satd4x4_abs(const typename Px<P>::Row4 f[4], const typename Px<P>::Row4 r[4])
{
    int t[4][4];
#pragma unroll
    for (int y = 0; y < 4; y++)
        had_row(f[y], r[y], t[y]);
    int sum = 0;
#pragma unroll
    for (int x = 0; x < 4; x++)
    {
        int s01 = t[0][x] + t[1][x], d01 = t[0][x] - t[1][x];
        int s23 = t[2][x] + t[3][x], d23 = t[2][x] - t[3][x];
        sum += abs(s01 + s23) + abs(d01 + d23) + abs(s01 - s23) + abs(d01 - d23);
    }
    return sum;
}

__device__ __forceinline__ int quad_sum(int v)
{
    v += __shfl_xor_sync(FULL_MASK, v, 1);
    v += __shfl_xor_sync(FULL_MASK, v, 2);
    return v;
}

__device__ __forceinline__ int warp_sum(int v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

#endif /* X265CU_DEV_CUH */
