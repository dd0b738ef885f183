"""Multi-rank check of the sample-sharded render on real GPUs (launched by torchrun, one rank per GPU; used by
tests/test_gpu_multi.py and by hand: python -m torch.distributed.run --nproc-per-node 2 scripts/multi_rank_check.py).

Every rank renders its shard of the samples of every pixel (sharding.shard_samples), ONE NCCL all-reduce sums the
radiance buffers, and rank 0 compares the result with its own single-GPU render of all samples:
  * the per-(pixel, sample) streams do not depend on the rank count, so the set of paths is identical: equal segment
    totals, equal NaN-poisoned pixels;
  * the sums differ only by the order of the float additions: relative difference <= 1e-5.
Then the view axis: a list of views dealt to the ranks in blocks, gathered, compared bit for bit with one rank rendering
them all.  Prints two JSON lines (rank 0)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import raytracingtherestofyourlife_b200 as B  # noqa: E402
from raytracingtherestofyourlife_b200.sharding import gather_views, render_views_sharded, shard_samples  # noqa: E402

W = int(os.environ.get("B2PT_CHECK_W", "512"))
SPP = int(os.environ.get("B2PT_CHECK_SPP", "48"))
DEPTH = int(os.environ.get("B2PT_CHECK_DEPTH", "50"))

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream(device=local)
torch.cuda.set_stream(stream)
ctx = B.Context(local)
ctx.set_stream(stream.cuda_stream)
ctx.set_scene(B.Scene.cornell())
ctx.build_bvh()
ctx.set_camera(B.Camera(W, W))
color = torch.zeros((W * W, 4), dtype=torch.float32, device="cuda")
ctx.set_color_tensor(color)
begin, count = shard_samples(SPP, rank, world)
ctx.clear_color()
ctx.render_range(begin, count, DEPTH, 0)
seg = torch.tensor([float(ctx.stats().segments)], dtype=torch.float64, device="cuda")
dist.all_reduce(color, op=dist.ReduceOp.SUM)
dist.all_reduce(seg, op=dist.ReduceOp.SUM)
torch.cuda.synchronize()
sharded = color.cpu().numpy().copy()
if rank == 0:
    ctx.clear_color()
    ctx.render_range(0, SPP, DEPTH, 0)
    torch.cuda.synchronize()
    single = color.cpu().numpy()
    seg1 = ctx.stats().segments
    nan_s, nan_1 = np.isnan(sharded[:, :3]), np.isnan(single[:, :3])
    ok = ~(nan_s | nan_1)
    rel = np.abs(sharded[:, :3][ok] - single[:, :3][ok]) / np.maximum(np.abs(single[:, :3][ok]), 1e-3)
    print(json.dumps({"world": world, "canvas": W, "spp": SPP, "depth": DEPTH, "segments_sharded": int(seg.item()),
                      "segments_single": int(seg1), "nan_masks_equal": bool(np.array_equal(nan_s, nan_1)),
                      "nan_channels": int(nan_1.sum()), "max_rel_diff": float(rel.max()),
                      "bit_identical_fraction": float((sharded[:, :3][ok] == single[:, :3][ok]).mean())}), flush=True)
# ---- the second axis: a list of views dealt to the ranks in contiguous blocks (sharding.render_views_sharded), no
# collective on the data path; the optional gather returns the whole stack in view order, bit-identical to one rank
# rendering every view
NV, VW, VSPP, VDEPTH = 11, 96, 12, 8
c = 278 / 555.0
views = np.array([[c + 2.2 * np.cos(t), c + 0.2 * np.sin(2 * t), c + 2.2 * np.sin(t), c, c, c, 0, 1, 0, 40.0]
                  for t in np.linspace(0.3, 5.9, NV)], np.float32)
begin_v, mine = render_views_sharded(lambda block: torch.from_numpy(
    ctx.render_views(block, VW, VW, VSPP, VDEPTH) if len(block) else np.zeros((0, VW * VW, 4), np.float32)).cuda(),
    views, rank, world)


def _all_gather(block):
    parts = [torch.empty_like(block) for _ in range(world)]
    dist.all_gather(parts, block)
    return parts


stack = gather_views(_all_gather, mine, NV, rank, world).cpu().numpy()
if rank == 0:
    whole = ctx.render_views(views, VW, VW, VSPP, VDEPTH)
    print(json.dumps({"world": world, "views": NV, "view_canvas": VW, "view_stack_bit_identical":
                      bool(np.array_equal(stack.view(np.uint32), np.ascontiguousarray(whole).view(np.uint32)))}), flush=True)
ctx.close()
dist.barrier()
dist.destroy_process_group()
