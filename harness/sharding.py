"""Multi-GPU partitioning of the lookahead path: independent encoder streams are the unit
(SURVEY.md §8e).  Streams are dealt round-robin to ranks; there is NO collective on the cost path.
The only exchange is the final gather of per-stream results (frames processed, checksum of
checksums) to rank 0, which is what a job scheduler needs to account throughput."""
import zlib


def assign(n_streams, world, rank):
    """stream ids owned by `rank` (round-robin: balanced to within one stream)"""
    return [s for s in range(n_streams) if s % world == rank]


def stream_seed(base_seed, stream_id):
    """every stream encodes its own clip"""
    return (base_seed + 7919 * stream_id) & 0x7FFFFFFF


def digest(values):
    """order-independent-free checksum of a list of ints (a checksum of checksums)"""
    return zlib.crc32(",".join(str(int(v)) for v in values).encode()) & 0xFFFFFFFF


def gather_results(dist, local):
    """all ranks contribute {stream_id: (frames, digest)}; returns the merged dict on every rank"""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    if world == 1:
        return dict(local)
    parts = [None] * world
    dist.all_gather_object(parts, dict(local))
    merged = {}
    for p in parts:
        for k, v in p.items():
            if k in merged:
                raise RuntimeError("stream %r processed twice" % (k,))
            merged[k] = v
    return merged
