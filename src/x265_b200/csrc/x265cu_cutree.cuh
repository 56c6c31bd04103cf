/* x265cu_cutree.cuh -- cuTree propagation on the device.
 *
 * Replaces Lookahead::estimateCUPropagate (encoder/slicetype.cpp:1741-1839) with its primitive
 * estimateCUPropagateCost (common/pixel.cpp:848-874) and the memsets of Lookahead::cuTree (:1668-1701).
 * It consumes exactly the arrays the estimate kernels left in HBM (intraCost, invQscaleFactor,
 * lowresCosts[d0][d1], lowresMvs[l][d]); nothing is uploaded for it.
 *
 * A cuTree pass is a CHAIN of steps: step k reads the propagateCost of frame b that steps < k accumulated and
 * scatters into the propagateCost of b's references.  One launch runs the whole chain: a thread-block cluster of
 * CUTREE_CTAS CTAs (8, the portable maximum), every step spread over all of its threads, the hardware cluster
 * barrier between steps (no grid-wide software barrier, no launch per step: a step is only nCU = 8160 / 32400
 * small work items, i.e. launch-latency bound on its own).  Steps that do not depend on each other share a PHASE
 * (no barrier between them): zeroing frames nobody is touching, and propagate steps that only ADD into common
 * frames (adds commute); the host marks the phase ends (cutree_mark_phases below).
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

#define CUTREE_CTAS 8
#define CUTREE_THREADS 1024
#define CUTREE_MAX_OPS 64      /* keeps the kernel parameter block under 4 KB */

enum { CT_OP_ZERO = 0, CT_OP_PROPAGATE = 1, CT_OP_PACK = 2 };

struct CutreeOpDev
{
    int kind;
    int fenc, ref0, ref1;       /* frame slots (ZERO / PACK: fenc only) */
    int costOfs;                /* (d0 * (bf + 2) + d1): which lowresCosts table of the frame */
    int mvOfs0, mvOfs1;         /* (list * (bf + 1) + d - 1): which MV field; -1: list not used by this estimate */
    int referenced;
    int bipredWeight;           /* bipredWeights[0]; [1] = 64 - it */
    int outIndex;               /* PACK: which staging area */
    int barrierAfter;           /* the next op depends on this phase: cluster barrier before it */
    double fps;                 /* fpsFactor / 256 (exact: power of two) */
};

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

__device__ __forceinline__ void cutree_clip_add(unsigned long long* cell, int x)
{
    if (x <= 0) return;                               /* adding 0 changes nothing; negatives cannot occur for legal inputs */
    atomicAdd(cell, (unsigned long long)(x < 65535 ? x : 65535));
}

__global__ void __launch_bounds__(CUTREE_THREADS, 1) cutree_kernel(const __grid_constant__ CutreeArgs a)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    const int tid = blockIdx.x * CUTREE_THREADS + threadIdx.x;
    const int nThreads = gridDim.x * CUTREE_THREADS;
    const int wCU = a.wCU, hCU = a.hCU, nCU = a.nCU;

    for (int k = 0; k < a.nOps; k++)
    {
        const CutreeOpDev& op = a.ops[k];
        unsigned long long* own = a.acc + (size_t)op.fenc * nCU;
        if (op.kind == CT_OP_ZERO)
        {
            for (int cu = tid; cu < nCU; cu += nThreads) __stcg(own + cu, 0ull);
        }
        else if (op.kind == CT_OP_PACK)
        {
            uint16_t* o = a.out + (size_t)op.outIndex * nCU;
            for (int cu = tid; cu < nCU; cu += nThreads)
            {
                const unsigned long long v = __ldcg(own + cu);
                o[cu] = (uint16_t)(v < 65535ull ? v : 65535ull);
            }
        }
        else
        {
            const int* intraCost = a.intraCost + (size_t)op.fenc * nCU;
            const int* invQ = a.invQ + (size_t)op.fenc * nCU;
            const uint16_t* costs = a.lowresCosts + ((size_t)op.fenc * a.costTables + op.costOfs) * nCU;
            unsigned long long* refAcc[2] = { a.acc + (size_t)op.ref0 * nCU, a.acc + (size_t)op.ref1 * nCU };
            const int* mvField[2] = { op.mvOfs0 >= 0 ? a.mvs + ((size_t)op.fenc * a.mvFields + op.mvOfs0) * nCU : NULL,
                                      op.mvOfs1 >= 0 ? a.mvs + ((size_t)op.fenc * a.mvFields + op.mvOfs1) * nCU : NULL };
            const int bw[2] = { op.bipredWeight, 64 - op.bipredWeight };
            for (int cu = tid; cu < nCU; cu += nThreads)
            {
                unsigned in = 0;
                if (op.referenced)
                {
                    const unsigned long long v = __ldcg(own + cu);
                    in = (unsigned)(v < 65535ull ? v : 65535ull);
                }
                const int cost = costs[cu];
                const int amount = cutree_amount(intraCost[cu], cost, invQ[cu], in, op.fps);
                if (amount <= 0) continue;            /* "don't propagate for an intra block" */
                const int listsUsed = cost >> 14;
                const int blocky = cu / wCU, blockx = cu - blocky * wCU;
#pragma unroll
                for (int list = 0; list < 2; list++)
                {
                    if (!((listsUsed >> list) & 1) || !mvField[list]) continue;
                    int listamount = amount;
                    if (listsUsed == 3)
                        listamount = (listamount * bw[list] + 32) >> 6;
                    const int mv = mvField[list][cu];
                    unsigned long long* ref = refAcc[list];
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
            /* "for non-referred frames the source costs are always zero, so just memset one row and re-use it" (:1757):
             * the first row of b's own array is zeroed by the step */
            if (!op.referenced)
                for (int cu = tid; cu < wCU; cu += nThreads) __stcg(own + cu, 0ull);
        }
        /* the next step reads (or must be ordered after) what this phase accumulated: make the atomics/stores
         * visible cluster-wide.  barrierAfter is uniform over the cluster (kernel parameter). */
        if (op.barrierAfter)
        {
            __threadfence();
            cluster.sync();
        }
    }
}

/* Host side: which ops may share a phase.  Per op, the frame arrays it READS, ADDS into (atomic, commutative) and
 * WRITES (plain stores).  Op k+1 joins the phase of op k unless it reads or writes an array the phase adds into or
 * writes, or adds into / writes an array the phase reads or writes (two ops adding into the same array are fine). */
static inline void cutree_mark_phases(CutreeOpDev* ops, int n)
{
    enum { R = 1, A = 2, W = 4 };
    struct Use { int slot, how; };
    Use phase[4 * CUTREE_MAX_OPS];
    int nPhase = 0;
    for (int k = 0; k < n; k++)
    {
        Use u[3];
        int nu = 0;
        const CutreeOpDev& o = ops[k];
        if (o.kind == CT_OP_ZERO) { u[nu].slot = o.fenc; u[nu++].how = W; }
        else if (o.kind == CT_OP_PACK) { u[nu].slot = o.fenc; u[nu++].how = R; }
        else
        {
            u[nu].slot = o.fenc; u[nu++].how = o.referenced ? R : W;      /* non-referenced: its first row is zeroed */
            u[nu].slot = o.ref0; u[nu++].how = A;
            if (o.mvOfs1 >= 0) { u[nu].slot = o.ref1; u[nu++].how = A; }
        }
        bool conflict = false;
        for (int i = 0; i < nu && !conflict; i++)
            for (int j = 0; j < nPhase && !conflict; j++)
                if (phase[j].slot == u[i].slot && !(phase[j].how == A && u[i].how == A) && !(phase[j].how == R && u[i].how == R))
                    conflict = true;
        if (conflict)
        {
            ops[k - 1].barrierAfter = 1;
            nPhase = 0;
        }
        for (int i = 0; i < nu; i++) phase[nPhase++] = u[i];
        ops[k].barrierAfter = 0;
    }
    if (n) ops[n - 1].barrierAfter = 1;    /* the launch ends a phase (kernel boundary orders the rest) */
}

#endif /* X265CU_CUTREE_CUH */
