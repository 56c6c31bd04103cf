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
 * The per-CU work itself lives in x265cu_cutree_core.h, compiled for the device here and for the CPU by
 * tests/cutree_emul.cpp (emulation against the oracle in the CPU test suite), like la_core.h for the motion search.
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

#include "x265cu_cutree_core.h"

#define CUTREE_THREADS 512
/* memory operations of the device: the accumulators live in L2 (no L1 copies survive a grid barrier), adds are 64-bit
 * reductions */
struct CutreeDeviceMem
{
    static __device__ __forceinline__ unsigned long long load(const unsigned long long* p) { return __ldcg(p); }
    static __device__ __forceinline__ void store(unsigned long long* p, unsigned long long v) { __stcg(p, v); }
    static __device__ __forceinline__ void add(unsigned long long* p, unsigned long long v) { atomicAdd(p, v); }
};

__device__ __forceinline__ void cutree_prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" :: "l"(p));
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
            cutree_item<CutreeDeviceMem>(a, a.ops[k0 + k], item - k * nCU);
        }
        k0 = k1;
        if (k0 < a.nOps) grid.sync();
    }
}

#endif /* X265CU_CUTREE_CUH */
