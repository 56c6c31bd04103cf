/* x265cu_cutree_sched.h -- host-side part of the cuTree kernel's interface: the op record and the scheduler that
 * orders the ops of one launch into phases of mutually independent ops.  Plain C++ (no CUDA), so that the scheduler
 * can be tested on the CPU (tests/cutree_sched_test.cpp); included by x265cu_cutree.cuh. */
#ifndef X265CU_CUTREE_SCHED_H
#define X265CU_CUTREE_SCHED_H

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
    int barrierAfter;           /* last op of its phase: grid barrier before the next op */
    double fps;                 /* fpsFactor / 256 (exact: power of two) */
};

/* Host side: schedule the ops of one launch into PHASES.  Per op, the frame arrays it READS, ADDS into (atomic,
 * commutative) and WRITES (plain stores).  Two ops conflict when they use a common array other than both adding
 * into it or both reading it; conflicting ops keep their order, everything else may move.  Each op gets the level
 * 1 + max(level of the earlier ops it conflicts with) (as-soon-as-possible schedule), ops are sorted by level
 * (stably), one phase per level.  A cuTree pass over k mini-GOPs is then not 6k steps long but about k + 2: the
 * zeroing comes first, the non-referenced B frames (which read no accumulator) next, and only the chain
 * P <- P <- P ... through the reference frames stays serial. */
static inline void cutree_schedule(CutreeOpDev* ops, int n)
{
    enum { R = 1, A = 2, W = 4 };
    struct Use { int slot, how; };
    Use use[CUTREE_MAX_OPS][3];
    int nUse[CUTREE_MAX_OPS], level[CUTREE_MAX_OPS];
    for (int k = 0; k < n; k++)
    {
        const CutreeOpDev& o = ops[k];
        int nu = 0;
        if (o.kind == CT_OP_ZERO) { use[k][nu].slot = o.fenc; use[k][nu++].how = W; }
        else if (o.kind == CT_OP_PACK) { use[k][nu].slot = o.fenc; use[k][nu++].how = R; }
        else
        {
            use[k][nu].slot = o.fenc; use[k][nu++].how = o.referenced ? R : W;    /* non-referenced: its first row is zeroed */
            use[k][nu].slot = o.ref0; use[k][nu++].how = A;
            if (o.mvOfs1 >= 0) { use[k][nu].slot = o.ref1; use[k][nu++].how = A; }
        }
        nUse[k] = nu;
        level[k] = 0;
        for (int e = 0; e < k; e++)
        {
            bool conflict = false;
            for (int i = 0; i < nu && !conflict; i++)
                for (int j = 0; j < nUse[e] && !conflict; j++)
                    if (use[e][j].slot == use[k][i].slot && !(use[e][j].how == A && use[k][i].how == A) && !(use[e][j].how == R && use[k][i].how == R))
                        conflict = true;
            if (conflict && level[e] + 1 > level[k]) level[k] = level[e] + 1;
        }
    }
    CutreeOpDev sorted[CUTREE_MAX_OPS];
    int m = 0, maxLevel = 0;
    for (int k = 0; k < n; k++) if (level[k] > maxLevel) maxLevel = level[k];
    for (int l = 0; l <= maxLevel; l++)
    {
        const int first = m;
        for (int k = 0; k < n; k++)
            if (level[k] == l) { sorted[m] = ops[k]; sorted[m].barrierAfter = 0; m++; }
        if (m > first) sorted[m - 1].barrierAfter = 1;
    }
    for (int k = 0; k < n; k++) ops[k] = sorted[k];
}

#endif /* X265CU_CUTREE_SCHED_H */
