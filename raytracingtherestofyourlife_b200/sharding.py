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


# ---- second axis: camera views.  A hemisphere sweep is a list of independent renders, so the
# views are dealt to the ranks in contiguous blocks and NO collective is needed on the data path: every rank owns
# (and typically writes out) the images of its own views.  gather_views is the optional convenience for callers that
# want the whole stack on every rank.
def shard_views(n_views, rank, world):
    """Contiguous [begin, begin+count) of the view list for `rank`; counts differ by at most one."""
    return shard_samples(n_views, rank, world)


def render_views_sharded(render_views, views, rank, world):
    """render_views(view_block) -> [len(view_block), ...] images of that block (Context.render_views on the GPU box).
    Returns (begin, images_of_this_rank)."""
    begin, count = shard_views(len(views), rank, world)
    return begin, render_views(views[begin:begin + count])


def gather_views(all_gather, images, n_views, rank, world):
    """Stack of all ranks' view images in view order.  all_gather(padded_block) -> list of `world` blocks (e.g. a thin
    wrapper over torch.distributed.all_gather); blocks are padded to the largest shard so that they have one shape."""
    import torch
    counts = [shard_views(n_views, r, world)[1] for r in range(world)]
    width = max(counts) if counts else 0
    block = torch.zeros((width,) + tuple(images.shape[1:]), dtype=images.dtype, device=images.device)
    block[:images.shape[0]] = images
    parts = all_gather(block)
    return torch.cat([parts[r][:counts[r]] for r in range(world)], dim=0)
