/* x265cu_cutree_core.h -- the per-CU work of the cuTree kernel (x265cu_cutree.cuh), written once for the device and
 * for its CPU emulation (tests/cutree_emul.cpp), like la_core.h for the motion search: the same source, the memory
 * operations supplied by a policy class (device: L2 loads/stores + 64-bit atomics; host: plain memory).
 *
 * Replaces Lookahead::estimateCUPropagate's CU loop (encoder/slicetype.cpp:1763-1835) and estimateCUPropagateCost
 * (common/pixel.cpp:848-874).  See x265cu_cutree.cuh for why the results are bit-identical to the reference. */
#ifndef X265CU_CUTREE_CORE_H
#define X265CU_CUTREE_CORE_H

#include <stdint.h>
#include <stddef.h>

#include "x265cu_cutree_sched.h"

#ifdef __CUDACC__
#define CT_HD __host__ __device__ __forceinline__
#else
#define CT_HD static inline
#endif

struct CutreeArgs
{
    int nOps;
    int wCU, hCU, nCU;
    int costTables;             /* (bf + 2)^2 */
    int mvFields;               /* 2 * (bf + 1) */
    const int* intraCost;       /* [slot][nCU] */
    const int* invQ;            /* [slot][nCU] */
    const uint16_t* lowresCosts;/* [slot][costTables][nCU] */
    const int* mvs;             /* [slot][mvFields][nCU] packed int16 x | int16 y << 16 */
    unsigned long long* acc;    /* [slot][nCU] propagateCost accumulators */
    uint16_t* out;              /* [outIndex][nCU] clamped copies for the host */
    CutreeOpDev ops[CUTREE_MAX_OPS];
};

/* one IEEE-754 round-to-nearest operation each, never contracted into an FMA */
#ifdef __CUDA_ARCH__
#define CT_DMUL(a, b) __dmul_rn(a, b)
#define CT_DADD(a, b) __dadd_rn(a, b)
#define CT_DDIV(a, b) __ddiv_rn(a, b)
#define CT_D2I(r) __double2int_rz(r)
#else
static inline double ct_vol(double v) { volatile double r = v; return r; }
#define CT_DMUL(a, b) ct_vol((a) * (b))
#define CT_DADD(a, b) ct_vol((a) + (b))
#define CT_DDIV(a, b) ct_vol((a) / (b))
#define CT_D2I(r) ((int)(r))
#endif

/* estimateCUPropagateCost for one CU */
CT_HD int cutree_amount(int intra, int interRaw, int invQ, unsigned in, double fps)
{
    int inter = interRaw & 0x3FFF;                    /* LOWRES_COST_MASK */
    inter = inter < intra ? inter : intra;
    const int prod = (int)((unsigned)intra * (unsigned)invQ);
    double r = CT_DMUL((double)prod, fps);
    r = CT_DADD(r, (double)in);
    r = CT_DMUL(r, (double)(intra - inter));
    r = CT_DDIV(r, (double)intra);
    r = CT_DADD(r, 0.5);
    /* cvttsd2si: NaN and out-of-range give INT_MIN (never propagated: the caller tests > 0) */
    return (r >= -2147483648.0 && r < 2147483648.0) ? CT_D2I(r) : (int)0x80000000;
}

/* CLIP_ADD: every addend is >= 0, so saturating adds in any order equal min(sum, 65535): clamp the addend, add, clamp on read */
template <class M>
CT_HD void cutree_clip_add(unsigned long long* cell, int x)
{
    if (x <= 0) return;                               /* adding 0 changes nothing; negatives cannot occur for legal inputs */
    M::add(cell, (unsigned long long)(x < 65535 ? x : 65535));
}

/* one work item = one CU of one op */
template <class M>
CT_HD void cutree_item(const CutreeArgs& a, const CutreeOpDev& op, int cu)
{
    const int wCU = a.wCU, hCU = a.hCU, nCU = a.nCU;
    unsigned long long* own = a.acc + (size_t)op.fenc * nCU;
    if (op.kind == CT_OP_ZERO)
    {
        M::store(own + cu, 0ull);
        return;
    }
    if (op.kind == CT_OP_PACK)
    {
        const unsigned long long v = M::load(own + cu);
        a.out[(size_t)op.outIndex * nCU + cu] = (uint16_t)(v < 65535ull ? v : 65535ull);
        return;
    }
    const size_t f = (size_t)op.fenc;
    unsigned in = 0;
    if (op.referenced)
    {
        const unsigned long long v = M::load(own + cu);
        in = (unsigned)(v < 65535ull ? v : 65535ull);
    }
    /* every load of the item is issued before anything is consumed: one L2 round trip */
    const int cost = a.lowresCosts[(f * a.costTables + op.costOfs) * nCU + cu];
    const int mvBoth[2] = { a.mvs[(f * a.mvFields + op.mvOfs0) * nCU + cu],
                            op.mvOfs1 >= 0 ? a.mvs[(f * a.mvFields + op.mvOfs1) * nCU + cu] : 0 };
    const int amount = cutree_amount(a.intraCost[f * nCU + cu], cost, a.invQ[f * nCU + cu], in, op.fps);
    /* "for non-referred frames the source costs are always zero, so just memset one row and re-use it" (:1757): the
     * first row of b's own array is zeroed by the step (nothing reads or adds into it in this phase) */
    if (!op.referenced && cu < wCU) M::store(own + cu, 0ull);
    if (amount <= 0) return;                      /* "don't propagate for an intra block" */
    const int listsUsed = cost >> 14;
    const int blocky = cu / wCU, blockx = cu - blocky * wCU;
    const int bw[2] = { op.bipredWeight, 64 - op.bipredWeight };
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int list = 0; list < 2; list++)
    {
        if (!((listsUsed >> list) & 1) || (list && op.mvOfs1 < 0)) continue;
        int listamount = amount;
        if (listsUsed == 3)
            listamount = (listamount * bw[list] + 32) >> 6;
        const int mv = mvBoth[list];
        unsigned long long* ref = a.acc + (size_t)(list ? op.ref1 : op.ref0) * nCU;
        if (!mv)
        {
            cutree_clip_add<M>(ref + cu, listamount);
            continue;
        }
        int x = (short)(mv & 0xFFFF), y = mv >> 16;
        const int cux = (x >> 5) + blockx, cuy = (y >> 5) + blocky;
        x &= 31; y &= 31;
        const int idx0 = cux + cuy * wCU;
        const bool inX0 = cux >= 0 && cux < wCU, inX1 = cux + 1 >= 0 && cux + 1 < wCU;
        const bool inY0 = cuy >= 0 && cuy < hCU, inY1 = cuy + 1 >= 0 && cuy + 1 < hCU;
        if (inX0 && inY0) cutree_clip_add<M>(ref + idx0, (listamount * ((32 - y) * (32 - x)) + 512) >> 10);
        if (inX1 && inY0) cutree_clip_add<M>(ref + idx0 + 1, (listamount * ((32 - y) * x) + 512) >> 10);
        if (inX0 && inY1) cutree_clip_add<M>(ref + idx0 + wCU, (listamount * (y * (32 - x)) + 512) >> 10);
        if (inX1 && inY1) cutree_clip_add<M>(ref + idx0 + wCU + 1, (listamount * (y * x) + 512) >> 10);
    }
}

#endif /* X265CU_CUTREE_CORE_H */
