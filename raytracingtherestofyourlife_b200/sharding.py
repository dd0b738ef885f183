"""Sample sharding across ranks: rank g renders a contiguous range of the global sample
indices of EVERY pixel; the only exchange is one all-reduce (sum) of the W*H float4 radiance sums.

Because every (pixel, sample) path owns its RNG stream (state0 = pixel + seedOffset + sample*0x9E3779B9),
the union of the ranks' paths is exactly the single-GPU set of paths: the N-rank image equals the 1-rank
image up to float summation order.
"""


def shard_samples(spp, rank, world):
    """Contiguous [begin, begin+count) of the global samples for `rank`; counts differ by at most one."""
    if spp < 0 or world < 1 or not (0 <= rank < world):
        raise ValueError("bad shard request spp=%r rank=%r world=%r" % (spp, rank, world))
    base, rem = divmod(spp, world)
    begin = rank * base + min(rank, rem)
    count = base + (1 if rank < rem else 0)
    return begin, count


def render_sharded(render_range, allreduce_sum, spp, rank, world):
    """Drive one rank of a sample-sharded render.

    render_range(begin, count) -> buffer holding this rank's un-normalised radiance sum
    allreduce_sum(buffer)      -> in-place sum over ranks (torch.distributed.all_reduce on the GPU box)
    Returns the reduced buffer (identical on every rank).
    """
    begin, count = shard_samples(spp, rank, world)
    buf = render_range(begin, count)
    if world > 1:
        allreduce_sum(buf)
    return buf
