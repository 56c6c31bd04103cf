/* cutree_emul.cpp -- CPU emulation of the cuTree kernel's per-CU work (src/x265_b200/csrc/x265cu_cutree_core.h, the very
 * source the device compiles) against the oracle's restatement of Lookahead::estimateCUPropagate.
 *
 * TEST INFRASTRUCTURE (built and run by tests/test_cutree_emul.py; links the oracle object, which is itself pinned
 * against the reference on every golden trace).  Random adversarial frames -- saturating propagateCost, vectors leaving
 * the picture on every side, intra CUs, all list combinations, bipred weights, inverse qscales up to 16x -- go through
 *   (a) ola_estimate_cu_propagate / ola_cutree_zero on ola_frame structs, in the given order, and
 *   (b) cutree_item<HostMem> over flat arrays laid out like the device mirrors, ops scheduled into phases by
 *       cutree_schedule() exactly as x265cu_cutree_run does, the CUs of a phase visited in a scrambled order;
 * every propagateCost array must be identical afterwards. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "../oracle/x265la_oracle.h"
#include "../src/x265_b200/csrc/x265cu_cutree_core.h"

struct HostMem
{
    static unsigned long long load(const unsigned long long* p) { return *p; }
    static void store(unsigned long long* p, unsigned long long v) { *p = v; }
    static void add(unsigned long long* p, unsigned long long v) { *p += v; }
};

static unsigned rnd() { static unsigned long long s = 88172645463325252ull; s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (unsigned)(s >> 16); }

int main(int argc, char** argv)
{
    const int rounds = argc > 1 ? atoi(argv[1]) : 40;
    const int BF = 4, SLOTS = 7;
    long steps = 0;
    for (int r = 0; r < rounds; r++)
    {
        const int srcW = 64 + 16 * (int)(rnd() % 12), srcH = 64 + 16 * (int)(rnd() % 8);
        ola_frame* fr[SLOTS];
        for (int s = 0; s < SLOTS; s++) fr[s] = ola_frame_create(srcW, srcH, 96, 80, BF, 1);
        const int wCU = fr[0]->g.wCU, hCU = fr[0]->g.hCU, nCU = fr[0]->g.nCU;
        const int costTables = (BF + 2) * (BF + 2), mvFields = 2 * (BF + 1);
        std::vector<int> intra((size_t)SLOTS * nCU), invq((size_t)SLOTS * nCU), mvs((size_t)SLOTS * mvFields * nCU);
        std::vector<uint16_t> costs((size_t)SLOTS * costTables * nCU);
        std::vector<unsigned long long> acc((size_t)SLOTS * nCU);
        const int maxIntra = ORACLE_DEPTH > 8 ? 1 << 16 : 1 << 14;
        for (int s = 0; s < SLOTS; s++)
        {
            for (int i = 0; i < nCU; i++)
            {
                fr[s]->intraCost[i] = intra[(size_t)s * nCU + i] = 1 + (int)(rnd() % maxIntra);
                fr[s]->invQscale[i] = invq[(size_t)s * nCU + i] = (rnd() % 8 == 0) ? 4096 : 16 + (int)(rnd() % 2048);
                static const int sat[7] = { 0, 1, 7, 40000, 65000, 65534, 65535 };
                fr[s]->propagateCost[i] = (uint16_t)((r & 1) ? sat[rnd() % 7] : rnd() % 65536);
                acc[(size_t)s * nCU + i] = fr[s]->propagateCost[i];
            }
            for (int d0 = 1; d0 <= BF + 1; d0++)
                for (int d1 = 0; d1 <= BF + 1; d1++)
                    for (int i = 0; i < nCU; i++)
                    {
                        int c = (int)(rnd() % (1 << 14));
                        if (rnd() % 10 < 7) { int half = fr[s]->intraCost[i] / 2; if (c > half) c = half; }
                        const int lists = d1 ? (int)(rnd() % 4) : (int)(rnd() % 2);
                        fr[s]->lowresCosts[d0][d1][i] = costs[((size_t)s * costTables + d0 * (BF + 2) + d1) * nCU + i] = (uint16_t)(c | (lists << 14));
                    }
            for (int l = 0; l < 2; l++)
                for (int d = 1; d <= BF + 1; d++)
                    for (int i = 0; i < nCU; i++)
                    {
                        int x, y;
                        const unsigned k = rnd() % 10;
                        if (k < 2) x = y = 0;
                        else if (k < 5) { x = (int)(rnd() % 8001) - 4000; y = (int)(rnd() % 8001) - 4000; }
                        else { x = (int)(rnd() % 129) - 64; y = (int)(rnd() % 129) - 64; }
                        fr[s]->mvs[l][d - 1][i].x = (int16_t)x; fr[s]->mvs[l][d - 1][i].y = (int16_t)y;
                        mvs[((size_t)s * mvFields + l * (BF + 1) + d - 1) * nCU + i] = (int)((unsigned)(uint16_t)x | ((unsigned)(uint16_t)y << 16));
                    }
        }
        /* a random op list over the slots: slot index = picture order, so b - p0 / p1 - b are the distances */
        CutreeArgs a;
        memset(&a, 0, sizeof(a));
        a.wCU = wCU; a.hCU = hCU; a.nCU = nCU; a.costTables = costTables; a.mvFields = mvFields;
        a.intraCost = &intra[0]; a.invQ = &invq[0]; a.lowresCosts = &costs[0]; a.mvs = &mvs[0]; a.acc = &acc[0]; a.out = NULL;
        const int n = 8 + (int)(rnd() % 40);
        for (int k = 0; k < n; k++)
        {
            CutreeOpDev& o = a.ops[a.nOps++];
            memset(&o, 0, sizeof(o));
            if (rnd() % 6 == 0)
            {
                o.kind = CT_OP_ZERO; o.fenc = (int)(rnd() % SLOTS);
                ola_cutree_zero(fr[o.fenc]);
                continue;
            }
            int p0, b, p1;
            do { p0 = (int)(rnd() % SLOTS); b = p0 + 1 + (int)(rnd() % (BF + 1)); p1 = (rnd() & 1) ? b : b + 1 + (int)(rnd() % (BF + 1)); }
            while (b >= SLOTS || p1 >= SLOTS);
            const int d0 = b - p0, d1 = p1 - b, referenced = (int)(rnd() & 1), weighted = (int)(rnd() & 1);
            const int dsf = ((d0 << 8) + ((p1 - p0) >> 1)) / (p1 - p0);
            o.kind = CT_OP_PROPAGATE; o.fenc = b; o.ref0 = p0; o.ref1 = p1;
            o.costOfs = d0 * (BF + 2) + d1; o.mvOfs0 = d0 - 1; o.mvOfs1 = d1 > 0 ? (BF + 1) + d1 - 1 : -1;
            o.referenced = referenced; o.bipredWeight = weighted ? 64 - (dsf >> 2) : 32;
            o.fps = 1.0 * (1.0 / 256);
            ola_estimate_cu_propagate(fr[b], fr[p0], fr[p1], d0, d1, referenced, 1.0 / 30, 30, 1, weighted);
            steps++;
        }
        cutree_schedule(a.ops, a.nOps);
        for (int k0 = 0; k0 < a.nOps;)
        {
            int k1 = k0;
            while (!a.ops[k1].barrierAfter) k1++;
            k1++;
            const int total = (k1 - k0) * nCU;
            const int stride = 7919;             /* coprime with every total here: a scrambled visiting order of the phase's items */
            for (int t = 0, item = (int)(rnd() % total); t < total; t++, item = (int)(((long long)item + stride) % total))
                cutree_item<HostMem>(a, a.ops[k0 + item / nCU], item % nCU);
            k0 = k1;
        }
        for (int s = 0; s < SLOTS; s++)
            for (int i = 0; i < nCU; i++)
            {
                const unsigned long long v = acc[(size_t)s * nCU + i];
                const uint16_t got = (uint16_t)(v < 65535 ? v : 65535);
                if (got != fr[s]->propagateCost[i])
                {
                    printf("round %d (%dx%d CUs, %d ops): slot %d cu %d: emulation %u, oracle %u\n", r, wCU, hCU, a.nOps, s, i, got, fr[s]->propagateCost[i]);
                    return 1;
                }
            }
        for (int s = 0; s < SLOTS; s++) ola_frame_destroy(fr[s]);
    }
    printf("ok: %d rounds, %ld propagate steps\n", rounds, steps);
    return 0;
}
