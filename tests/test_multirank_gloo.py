"""world_size-2 gloo test of the sample-sharding host logic: each rank renders its sample
range, one all-reduce sums the radiance buffers, and the result equals the single-rank render up to float
summation order.  The per-rank renderer here is the CPU oracle's production-stream mode (the GPU path is
checked the same way on the GPU box, tests/test_gpu_multi.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, spp, depth, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from raytracingtherestofyourlife_b200.sharding import render_sharded
    sc, cam = O.cornell_scene(), O.Camera(48, 32)

    def render_range(begin, count):
        img, _ = O.render(sc, cam, count, depth, mode=O.MODE_FORWARD_FAST, sample_begin=begin, threads=2)
        return torch.from_numpy(img)

    buf = render_sharded(render_range, lambda t: dist.all_reduce(t, op=dist.ReduceOp.SUM), spp, rank, world)
    np.save(os.path.join(out_dir, "rank%d.npy" % rank), buf.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("spp", [7, 8])
def test_two_rank_sharded_render_equals_single_rank(tmp_path, oracle, spp):
    world, depth = 2, 6
    mp.spawn(_worker, args=(world, _free_port(), spp, depth, str(tmp_path)), nprocs=world, join=True)
    r0 = np.load(tmp_path / "rank0.npy")
    r1 = np.load(tmp_path / "rank1.npy")
    assert np.array_equal(r0, r1, equal_nan=True)  # all-reduce leaves identical buffers on every rank
    single, _ = oracle.render(oracle.cornell_scene(), oracle.Camera(48, 32), spp, depth, mode=oracle.MODE_FORWARD_FAST)
    ok = ~np.isnan(single)
    assert np.array_equal(np.isnan(single), np.isnan(r0))
    assert np.allclose(r0[ok], single[ok], rtol=1e-5, atol=1e-6)


def _views_worker(rank, world, port, n_views, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle as O
    from raytracingtherestofyourlife_b200.sharding import render_views_sharded, gather_views
    sc = O.cornell_scene()
    views = _test_views(n_views)

    def render_views(block):  # stands in for Context.render_views (one image stack per view block)
        imgs = [O.render(sc, O.Camera(24, 16, pos=v[0:3], lookAt=v[3:6], up=tuple(v[6:9]), fov=float(v[9])), 3, 4,
                         mode=O.MODE_FORWARD_FAST, threads=2)[0] for v in block]
        return torch.from_numpy(np.stack(imgs)) if imgs else torch.zeros((0, 24 * 16, 4))

    begin, mine = render_views_sharded(render_views, views, rank, world)

    def all_gather(block):
        parts = [torch.empty_like(block) for _ in range(world)]
        dist.all_gather(parts, block)
        return parts

    full = gather_views(all_gather, mine, n_views, rank, world)
    np.save(os.path.join(out_dir, "views_rank%d.npy" % rank), full.numpy())
    np.save(os.path.join(out_dir, "begin_rank%d.npy" % rank), np.array([begin, mine.shape[0]]))
    dist.barrier()
    dist.destroy_process_group()


def _test_views(n):
    c = np.array([278 / 555.0] * 3, np.float32)
    out = []
    for k in range(n):
        pos = c + np.array([0.3 * (k - n / 2.0), 0.1 * k, -1078 / 555.0], np.float32)
        out.append(np.concatenate([pos, c, [0, 1, 0], [40.0 + k]]).astype(np.float32))
    return np.stack(out)


@pytest.mark.parametrize("n_views", [5, 1])
def test_two_rank_view_sharding(tmp_path, oracle, n_views):
    """Views are independent renders: each rank renders a contiguous block of the list (no collective on the data
    path); the optional gather returns the whole stack, in view order, bit-identical to rendering them on one rank."""
    world = 2
    mp.spawn(_views_worker, args=(world, _free_port(), n_views, str(tmp_path)), nprocs=world, join=True)
    full0, full1 = np.load(tmp_path / "views_rank0.npy"), np.load(tmp_path / "views_rank1.npy")
    assert full0.shape == (n_views, 24 * 16, 4)
    assert np.array_equal(full0.view(np.uint32), full1.view(np.uint32))
    b0, b1 = np.load(tmp_path / "begin_rank0.npy"), np.load(tmp_path / "begin_rank1.npy")
    assert b0[0] == 0 and b1[0] == b0[1] and b0[1] + b1[1] == n_views
    sc = oracle.cornell_scene()
    for k, v in enumerate(_test_views(n_views)):
        ref, _ = oracle.render(sc, oracle.Camera(24, 16, pos=v[0:3], lookAt=v[3:6], up=tuple(v[6:9]), fov=float(v[9])),
                               3, 4, mode=oracle.MODE_FORWARD_FAST)
        assert np.array_equal(full0[k].view(np.uint32), ref.view(np.uint32)), k


def test_shard_samples_partition():
    from raytracingtherestofyourlife_b200.sharding import shard_samples
    for spp in (0, 1, 7, 8, 1024, 4096):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_samples(spp, r, world) for r in range(world)]
            assert sum(c for _, c in ranges) == spp
            pos = 0
            for b, c in ranges:
                assert b == pos and c >= 0
                pos += c
            assert max(c for _, c in ranges) - min(c for _, c in ranges) <= 1
    with pytest.raises(ValueError):
        shard_samples(8, 2, 2)
