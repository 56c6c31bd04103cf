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

template <typename P>
__device__ __forceinline__ int satd4x4_abs(const typename Px<P>::Row4 f[4], const typename Px<P>::Row4 r[4])
{
    int d[4][4];
#pragma unroll
    for (int y = 0; y < 4; y++)
    {
        int a[4], b[4];
        Px<P>::unpack(f[y], a);
        Px<P>::unpack(r[y], b);
#pragma unroll
        for (int x = 0; x < 4; x++) d[y][x] = a[x] - b[x];
    }
    return hadamard4x4_abs(d);
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
