"""N > 1 path on CPU: two ranks over gloo, streams sharded round-robin, no data-path collective;
each rank runs its streams' lookahead (here: the oracle on the tiny clip with a per-stream seed)
and the results are gathered once.  Checks the partition is exact and rank-independent."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_assign_is_a_partition():
    from harness import sharding
    for n in (1, 2, 7, 64):
        for world in (1, 2, 4, 8):
            seen = sorted(s for r in range(world) for s in sharding.assign(n, world, r))
            assert seen == list(range(n))
            sizes = [len(sharding.assign(n, world, r)) for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _stream_digest(stream_id):
    """run the lookahead of one stream on the CPU oracle and digest every checksum it produces"""
    from harness import sharding
    from oracle import pyoracle as po
    t = po.Trace(os.path.join(ROOT, "tests", "golden", "tiny8.trace"))
    t.cfg["seed"] = sharding.stream_seed(t.cfg["seed"], stream_id)
    t.events = [e for e in t.events if e[0] == "P"][:4]     # pre-lookahead of 4 frames is enough here
    r = po.OracleReplay(t)
    r.run()
    vals = []
    for poc in sorted(r.frames):
        f = r.frames[poc]
        vals += [po.crc(f.planes()), po.crc(f.intra_cost()), int(f.c.costEst[0][0])]
    r.close()
    return len(t.events), sharding.digest(vals)


def _worker(rank, world, port, n_streams, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from harness import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = {s: _stream_digest(s) for s in sharding.assign(n_streams, world, rank)}
    merged = sharding.gather_results(dist, mine)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, sorted(merged.items())))


def test_two_ranks_gloo():
    n_streams, world = 5, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    by_rank = dict(outs)
    assert by_rank[0] == by_rank[1]                      # every rank sees the same merged result
    assert [k for k, _ in by_rank[0]] == list(range(n_streams))
    single = sorted((s, _stream_digest(s)) for s in range(n_streams))
    assert by_rank[0] == single                          # identical to an unsharded run
    assert len(set(d for _, (_, d) in single)) == n_streams   # streams really differ
