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
