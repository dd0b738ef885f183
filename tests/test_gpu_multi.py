"""The N-rank GPU render against the one-rank render: launched through torchrun on the GPUs of the
box, skipped where fewer than two are visible (the driver's one-GPU box); `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_multi.py -m gpu` runs it.  Also the single-process form: G contexts in one process, sums added by
b2pt_allreduce (k_sum_peers over NVLink peer access), through the C-ABI and through the C++ facade."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_nccl_render_equals_single_gpu(b2pt, world):
    if _gpus() < world:
        pytest.skip("needs %d GPUs" % world)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                          "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                          os.path.join(ROOT, "scripts", "multi_rank_check.py")], capture_output=True, text=True, env=env,
                         timeout=900)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]
    line = [l for l in lines if "segments_sharded" in l][-1]
    assert line["world"] == world
    assert line["segments_sharded"] == line["segments_single"]  # the same set of paths
    assert line["nan_masks_equal"]
    assert line["max_rel_diff"] <= 1e-5  # summation order only
    vline = [l for l in lines if "view_stack_bit_identical" in l][-1]  # views dealt to the ranks: independent renders
    assert vline["world"] == world and vline["view_stack_bit_identical"]


def test_single_process_allreduce_over_peer_access(b2pt):
    """b2pt_allreduce: two contexts on two GPUs in THIS process, each renders half of the samples."""
    import ctypes as C
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    W, spp, depth = 256, 32, 50
    ctxs = [b2pt.Context(g) for g in range(2)]
    try:
        for g, ctx in enumerate(ctxs):
            ctx.set_scene(b2pt.Scene.cornell())
            ctx.build_bvh()
            ctx.set_camera(b2pt.Camera(W, W))
            ctx.clear_color()
            ctx.render_range(g * spp // 2, spp // 2, depth, 0)
        handles = (C.c_void_p * 2)(*[c._h for c in ctxs])
        rc = b2pt.lib().b2pt_allreduce(handles, 2)
        assert rc == 0, b2pt.lib().b2pt_last_error()
        a, b = ctxs[0].read_color().copy(), ctxs[1].read_color().copy()
        seg = sum(c.stats().segments for c in ctxs)
        ctxs[0].render(spp, depth, 0)
        one, seg1 = ctxs[0].read_color(), ctxs[0].stats().segments
    finally:
        for c in ctxs:
            c.close()
    assert np.array_equal(a, b, equal_nan=True)  # every context ends with the same sum
    assert seg == seg1
    assert np.array_equal(np.isnan(a), np.isnan(one))
    ok = ~np.isnan(one)
    assert (np.abs(a[ok] - one[ok]) <= 1e-5 * np.maximum(np.abs(one[ok]), 1e-3)).all()


def test_facade_set_devices(b2pt):
    """MapperPathTracer::SetDevices({0, 1}) (host/test_facade --multigpu 2)."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "raytracingtherestofyourlife_b200", "host", "test_facade")
    out = subprocess.run([exe, "--multigpu", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert "all facade checks passed" in out.stdout
