"""Frame-parallel sharding of a batch / sequence across the GPUs of one box.

Frames are independent inside ORBextractor::operator() (no state is carried between
calls: mvImagePyramid is overwritten, src/ORBextractor.cc:1119), so rank g of G takes
the contiguous block [g*F/G, (g+1)*F/G) and no data-path collective is needed.  Blocks
are contiguous so frame-to-frame matching pairs (i, i+1) stay on one rank except at the
G-1 seams; a seam pair is owned by the rank holding frame i and needs frame i+1's
descriptors from the next rank's host output (or a redundant extraction of one frame).
"""


def shard_bounds(nframes, rank, world):
    """[lo, hi) of the frames rank `rank` owns."""
    if world < 1 or not 0 <= rank < world or nframes < 0:
        raise ValueError("bad shard request")
    return nframes * rank // world, nframes * (rank + 1) // world


def seam_pairs(nframes, world):
    """Consecutive-frame pairs (i, i+1) whose frames live on different ranks."""
    out = []
    for r in range(world - 1):
        hi = shard_bounds(nframes, r, world)[1]
        if 0 < hi < nframes:
            out.append((hi - 1, hi))
    return out


def job_throughput(units_per_rank, seconds_per_rank):
    """Whole-job rate from per-rank unit counts and device times: all units / slowest rank."""
    return sum(units_per_rank) / max(seconds_per_rank)
