/* x265cu_cutree.cuh -- cuTree propagation on the device.
 *
 * Replaces Lookahead::estimateCUPropagate (encoder/slicetype.cpp:1741-1839) with its primitive
 * estimateCUPropagateCost (common/pixel.cpp:848-874) and the memsets of Lookahead::cuTree (:1668-1701).
 * It consumes exactly the arrays the estimate kernels left in HBM (intraCost, invQscaleFactor,
 * lowresCosts[d0][d1], lowresMvs[l][d]); nothing is uploaded for it.
 *
 * A cuTree pass is a CHAIN of steps: step k reads the propagateCost of frame b that steps < k accumulated and
 * scatters into the propagateCost of b's references; a step is only nCU = 8160 / 32400 small work items, so on its
 * own it is launch-latency bound.  One COOPERATIVE launch (one CTA per SM) runs a whole pass:
 *  - the host schedules the steps into PHASES of mutually independent steps (cutree_schedule below: zeroing first,
 *    then every non-referenced B frame, then the serial chain through the reference frames: about k + 2 phases for
 *    k mini-GOPs instead of 6k steps);
 *  - the CUs of all steps of a phase are one pool of work items spread over the whole grid;
 *  - a grid barrier separates the phases.
 * Measured on c1_1080p (16 passes of up to 60 steps per 60 frames): a launch per pass on an 8-CTA cluster with the
 * hardware cluster barrier between steps took 1.75 ms per 60 frames whatever the number of barriers -- it was bound
 * by the instruction throughput of its 8 SMs (about 250 instructions per CU: the double division, the index
 * division, eight predicated atomics); hence the whole-GPU grid.
 *
 * Exactness.
 *  - The propagate amount uses double arithmetic in the reference.  Its object code performs, per CU,
 *    cvt(int32 intra*invQ) * (fpsFactor/256) + cvt(in), * cvt(intra - min(intra, inter)), / cvt(intra), + 0.5,
 *    truncate -- every step one IEEE-754 round-to-nearest operation, no contraction.  The same sequence is
 *    issued here with __dmul_rn/__dadd_rn/__ddiv_rn (never fused), so results are bit-identical.
 *  - CLIP_ADD saturates a uint16 at 65535.  Every addend is >= 0, so a chain of saturating adds equals
 *    min(sum, 65535) whatever the order: the accumulators are 64-bit (addends clamped to 65535 first, so they
 *    cannot overflow), added with atomics in any order, and clamped whenever they are read.
 */
#ifndef X265CU_CUTREE_CUH
#define X265CU_CUTREE_CUH

#include <cooperative_groups.h>

#include "x265cu_cutree_sched.h"

#define CUTREE_THREADS 512
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

/* estimateCUPropagateCost for one CU */
__device__ __forceinline__ int cutree_amount(int intra, int interRaw, int invQ, unsigned in, double fps)
{
    int inter = interRaw & 0x3FFF;                    /* LOWRES_COST_MASK */
    inter = inter < intra ? inter : intra;
    const int prod = (int)((unsigned)intra * (unsigned)invQ);
    double r = __dmul_rn((double)prod, fps);
    r = __dadd_rn(r, (double)in);
    r = __dmul_rn(r, (double)(intra - inter));
    r = __ddiv_rn(r, (double)intra);
    r = __dadd_rn(r, 0.5);
    /* cvttsd2si: NaN and out-of-range give INT_MIN (never propagated: the caller tests > 0) */
    return (r >= -2147483648.0 && r < 2147483648.0) ? __double2int_rz(r) : (int)0x80000000;
}

__device__ __forceinline__ void cutree_prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
}

__device__ __forceinline__ void cutree_clip_add(unsigned long long* cell, int x)
{
    if (x <= 0) return;                               /* adding 0 changes nothing; negatives cannot occur for legal inputs */
    atomicAdd(cell, (unsigned long long)(x < 65535 ? x : 65535));
}

/* one work item = one CU of one op */
__device__ __forceinline__ void cutree_item(const CutreeArgs& a, const CutreeOpDev& op, int cu)
{
    const int wCU = a.wCU, hCU = a.hCU, nCU = a.nCU;
    unsigned long long* own = a.acc + (size_t)op.fenc * nCU;
    if (op.kind == CT_OP_ZERO)
    {
        __stcg(own + cu, 0ull);
        return;
    }
    if (op.kind == CT_OP_PACK)
    {
        const unsigned long long v = __ldcg(own + cu);
        a.out[(size_t)op.outIndex * nCU + cu] = (uint16_t)(v < 65535ull ? v : 65535ull);
        return;
    }
    const size_t f = (size_t)op.fenc;
    unsigned in = 0;
    if (op.referenced)
    {
        const unsigned long long v = __ldcg(own + cu);
        in = (unsigned)(v < 65535ull ? v : 65535ull);
    }
    /* every load of the item is issued before anything is consumed: one L2 round trip */
    const int cost = a.lowresCosts[(f * a.costTables + op.costOfs) * nCU + cu];
    const int mvBoth[2] = { a.mvs[(f * a.mvFields + op.mvOfs0) * nCU + cu],
                            op.mvOfs1 >= 0 ? a.mvs[(f * a.mvFields + op.mvOfs1) * nCU + cu] : 0 };
    const int amount = cutree_amount(a.intraCost[f * nCU + cu], cost, a.invQ[f * nCU + cu], in, op.fps);
    /* "for non-referred frames the source costs are always zero, so just memset one row and re-use it" (:1757): the
     * first row of b's own array is zeroed by the step (nothing reads or adds into it in this phase) */
    if (!op.referenced && cu < wCU) __stcg(own + cu, 0ull);
    if (amount <= 0) return;                      /* "don't propagate for an intra block" */
    const int listsUsed = cost >> 14;
    const int blocky = cu / wCU, blockx = cu - blocky * wCU;
    const int bw[2] = { op.bipredWeight, 64 - op.bipredWeight };
#pragma unroll
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
            cutree_clip_add(ref + cu, listamount);
            continue;
        }
        int x = (short)(mv & 0xFFFF), y = mv >> 16;
        const int cux = (x >> 5) + blockx, cuy = (y >> 5) + blocky;
        x &= 31; y &= 31;
        const int idx0 = cux + cuy * wCU;
        const bool inX0 = cux >= 0 && cux < wCU, inX1 = cux + 1 >= 0 && cux + 1 < wCU;
        const bool inY0 = cuy >= 0 && cuy < hCU, inY1 = cuy + 1 >= 0 && cuy + 1 < hCU;
        if (inX0 && inY0) cutree_clip_add(ref + idx0, (listamount * ((32 - y) * (32 - x)) + 512) >> 10);
        if (inX1 && inY0) cutree_clip_add(ref + idx0 + 1, (listamount * ((32 - y) * x) + 512) >> 10);
        if (inX0 && inY1) cutree_clip_add(ref + idx0 + wCU, (listamount * (y * (32 - x)) + 512) >> 10);
        if (inX1 && inY1) cutree_clip_add(ref + idx0 + wCU + 1, (listamount * (y * x) + 512) >> 10);
    }
}

__global__ void __launch_bounds__(CUTREE_THREADS, 1) cutree_kernel(const __grid_constant__ CutreeArgs a)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    const int tid = blockIdx.x * CUTREE_THREADS + threadIdx.x;
    const int nThreads = gridDim.x * CUTREE_THREADS;
    const int nCU = a.nCU;

    /* The chain only carries the accumulators; what a step reads besides them (costs, intra costs, inverse qscales,
     * vectors) was written by kernels long ago and sits in HBM.  Pull all of it into L2 up front, for every step of
     * the launch at once, so that no phase of the chain waits for DRAM. */
    for (int item = tid; item < a.nOps * nCU; item += nThreads)
    {
        const int k = item / nCU, cu = item - k * nCU;
        const CutreeOpDev& op = a.ops[k];
        if (op.kind != CT_OP_PROPAGATE || (cu & 7)) continue;      /* one prefetch per 32-byte sector of the 4-byte arrays */
        const size_t f = (size_t)op.fenc;
        cutree_prefetch_l2(a.lowresCosts + (f * a.costTables + op.costOfs) * nCU + cu);
        cutree_prefetch_l2(a.intraCost + f * nCU + cu);
        cutree_prefetch_l2(a.invQ + f * nCU + cu);
        cutree_prefetch_l2(a.mvs + (f * a.mvFields + op.mvOfs0) * nCU + cu);
        if (op.mvOfs1 >= 0) cutree_prefetch_l2(a.mvs + (f * a.mvFields + op.mvOfs1) * nCU + cu);
    }

    /* phases: ops [k0, k1) are independent of each other (cutree_schedule); their CUs are ONE pool of work items
     * spread over the whole grid; the next phase reads what this one accumulated, so a grid barrier (with its
     * fences) separates them.  Phase bounds come from the kernel parameters: uniform over the grid. */
    for (int k0 = 0; k0 < a.nOps;)
    {
        int k1 = k0;
        while (!a.ops[k1].barrierAfter) k1++;
        k1++;
        const int total = (k1 - k0) * nCU;
        for (int item = tid; item < total; item += nThreads)
        {
            const int k = item / nCU;
            cutree_item(a, a.ops[k0 + k], item - k * nCU);
        }
        k0 = k1;
        if (k0 < a.nOps) grid.sync();
    }
}

#endif /* X265CU_CUTREE_CUH */
