/* cutree_sched_test.cpp -- CPU check of the cuTree op scheduler (src/x265_b200/csrc/x265cu_cutree_sched.h).
 *
 * TEST INFRASTRUCTURE (built and run by tests/test_cutree_sched.py).  A toy model of the accumulators executes random
 * op lists (a) in the given order, one op after the other -- the semantics the reference defines -- and (b) as the
 * kernel does: ops sorted into phases by cutree_schedule(), the ops of a phase executed in an arbitrary order with
 * all of their READS taken before any of their effects (the kernel gives no order at all inside a phase, so the
 * result must not depend on it).  Both must leave identical accumulators and identical packed outputs. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../src/x265_b200/csrc/x265cu_cutree_sched.h"

enum { SLOTS = 9, CELLS = 6 };

struct State
{
    uint64_t acc[SLOTS][CELLS];
    uint64_t out[CUTREE_MAX_OPS][CELLS];
};

static uint64_t mix(uint64_t v, uint64_t k) { return (v * 0x9E3779B97F4A7C15ull + k * 0xD1B54A32D192ED03ull) >> 40; }

/* effect of one op given the values it READS (own[] snapshot) */
struct Effect { int kind; int slot[3]; uint64_t add[2][CELLS]; int outIndex; uint64_t packed[CELLS]; int zeroFirst; };

static Effect evaluate(const CutreeOpDev& o, const State& s, int id)
{
    Effect e;
    memset(&e, 0, sizeof(e));
    e.kind = o.kind;
    e.slot[0] = o.fenc; e.slot[1] = o.ref0; e.slot[2] = o.mvOfs1 >= 0 ? o.ref1 : -1;
    if (o.kind == CT_OP_PACK)
    {
        e.outIndex = o.outIndex;
        for (int c = 0; c < CELLS; c++) e.packed[c] = s.acc[o.fenc][c] < 65535 ? s.acc[o.fenc][c] : 65535;
    }
    else if (o.kind == CT_OP_PROPAGATE)
    {
        for (int c = 0; c < CELLS; c++)
        {
            uint64_t in = o.referenced ? (s.acc[o.fenc][c] < 65535 ? s.acc[o.fenc][c] : 65535) : 0;
            e.add[0][(c + id) % CELLS] += mix(in + 1, id) & 0xFFFF;
            if (e.slot[2] >= 0) e.add[1][(c + 2 * id + 1) % CELLS] += mix(in + 7, id + 100) & 0xFFFF;
        }
        e.zeroFirst = !o.referenced;
    }
    return e;
}

static void apply(const Effect& e, State& s)
{
    if (e.kind == CT_OP_ZERO) memset(s.acc[e.slot[0]], 0, sizeof(s.acc[0]));
    else if (e.kind == CT_OP_PACK) memcpy(s.out[e.outIndex], e.packed, sizeof(e.packed));
    else
    {
        for (int c = 0; c < CELLS; c++) { s.acc[e.slot[1]][c] += e.add[0][c]; if (e.slot[2] >= 0) s.acc[e.slot[2]][c] += e.add[1][c]; }
        if (e.zeroFirst) s.acc[e.slot[0]][0] = 0;
    }
}

int main(int argc, char** argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 20000;
    srand(12345);
    long phasesTotal = 0, opsTotal = 0;
    for (int r = 0; r < rounds; r++)
    {
        const int n = 1 + rand() % CUTREE_MAX_OPS;
        CutreeOpDev ops[CUTREE_MAX_OPS], sorted[CUTREE_MAX_OPS];
        int ids[CUTREE_MAX_OPS];
        int nOut = 0;
        for (int k = 0; k < n; k++)
        {
            CutreeOpDev& o = ops[k];
            memset(&o, 0, sizeof(o));
            const int t = rand() % 10;
            o.fenc = rand() % SLOTS;
            if (t < 2) o.kind = CT_OP_ZERO;
            else if (t < 3) { o.kind = CT_OP_PACK; o.outIndex = nOut++; }
            else
            {
                o.kind = CT_OP_PROPAGATE;
                do { o.ref0 = rand() % SLOTS; } while (o.ref0 == o.fenc);
                if (rand() & 1) { do { o.ref1 = rand() % SLOTS; } while (o.ref1 == o.fenc); o.mvOfs1 = 3; }   /* ref1 may equal ref0 */
                else { o.ref1 = o.fenc; o.mvOfs1 = -1; }                                                     /* P-type: ref1 = own, unused */
                o.referenced = rand() & 1;
            }
            o.costOfs = k;      /* carries the op's identity through the sort */
        }
        State a, b;
        memset(&a, 0, sizeof(a));
        for (int s = 0; s < SLOTS; s++) for (int c = 0; c < CELLS; c++) a.acc[s][c] = rand() % 70000;
        b = a;
        for (int k = 0; k < n; k++) apply(evaluate(ops[k], a, k), a);          /* (a) the given order */

        memcpy(sorted, ops, sizeof(ops));
        cutree_schedule(sorted, n);
        if (!sorted[n - 1].barrierAfter) { printf("round %d: last op does not end a phase\n", r); return 1; }
        for (int k0 = 0; k0 < n;)
        {
            int k1 = k0;
            while (!sorted[k1].barrierAfter) k1++;
            k1++;
            /* (b) a phase: every op reads the state as it was when the phase began or as any subset of the phase left
             * it -- both extremes are tried: reads first, then effects in a shuffled order */
            std::vector<Effect> eff;
            for (int k = k0; k < k1; k++) { ids[k] = sorted[k].costOfs; eff.push_back(evaluate(sorted[k], b, ids[k])); }
            for (size_t i = eff.size(); i > 1; i--) { size_t j = (size_t)rand() % i; Effect t = eff[i - 1]; eff[i - 1] = eff[j]; eff[j] = t; }
            for (size_t i = 0; i < eff.size(); i++) apply(eff[i], b);
            phasesTotal++;
            k0 = k1;
        }
        opsTotal += n;
        if (memcmp(&a, &b, sizeof(a)))
        {
            printf("round %d (%d ops): scheduled execution differs from the given order\n", r, n);
            return 1;
        }
    }
    printf("ok: %d op lists, %ld ops in %ld phases\n", rounds, opsTotal, phasesTotal);
    return 0;
}
