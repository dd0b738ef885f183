"""The one-kernel-per-bounce pipeline of small scenes (B2PT_FLAG_ONE_KERNEL_BOUNCE; k_bounce: shade the hits of bounce
d-1, trace bounce d, bin the new hits; no ray queue) against the default two-kernel pipeline (k_trace + k_shade), which
the parity tests pin to the oracle: the same paths, the same per-path arithmetic, the same sample-order sums -- images must agree
bit for bit, in every mode that changes how records are laid out."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def same(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def render(ctx, spp, depth, flags):
    ctx.render(spp, depth, flags)
    return ctx.read_color().copy(), ctx.stats()


@pytest.mark.parametrize("W,H,spp,depth", [(128, 128, 10, 5), (96, 64, 48, 50), (64, 64, 64, 1), (50, 50, 32, 2),
                                           (33, 7, 40, 3), (256, 256, 32, 16)])
def test_one_kernel_pipeline_equals_two_kernel_pipeline(gpu_ctx, b2pt, W, H, spp, depth):
    gpu_ctx.set_camera(b2pt.Camera(W, H))
    a, sa = render(gpu_ctx, spp, depth, b2pt.FLAG_ONE_KERNEL_BOUNCE)
    b, sb = render(gpu_ctx, spp, depth, 0)
    assert sa.segments == sb.segments and sa.nanSamples == sb.nanSamples and sa.paths == sb.paths
    assert same(a, b)
    # one launch per bounce + the closing pass (+ candidate-mask prep when the canvas allows it + accumulate)
    assert sa.launches <= depth + 3 and sb.launches >= 2 * depth + 1


def test_one_kernel_pipeline_modes(gpu_ctx, b2pt, monkeypatch):
    """Several batches in flight, tail mode, the persistent cluster loop, reference-stream mode, the zero-throughput
    kill and the opt-out flags: every combination equals the two-kernel pipeline."""
    W = 192
    gpu_ctx.set_camera(b2pt.Camera(W, W))
    monkeypatch.setenv("B2PT_BATCH_PATHS", str(W * W * 4))
    for flags in (0, b2pt.FLAG_NO_TAIL, b2pt.FLAG_NO_OVERLAP, b2pt.FLAG_KILL_ZERO_THROUGHPUT, b2pt.FLAG_NO_AA,
                  b2pt.FLAG_NO_DEDUP, b2pt.FLAG_NO_PRIMARY_MASKS):
        a, sa = render(gpu_ctx, 24, 50, flags)
        b, sb = render(gpu_ctx, 24, 50, flags | b2pt.FLAG_ONE_KERNEL_BOUNCE)
        assert sa.batches == sb.batches == 6 and sa.segments == sb.segments, flags
        assert same(a, b), flags
    for per_warp, loop_rays in (("100000", "24576"), ("100000", "100000000"), ("256", "0")):
        monkeypatch.setenv("B2PT_TAIL_RAYS_PER_WARP", per_warp)
        monkeypatch.setenv("B2PT_TAIL_LOOP_RAYS", loop_rays)
        for depth in (50, 3, 2):
            a, sa = render(gpu_ctx, 24, depth, 0)
            b, sb = render(gpu_ctx, 24, depth, b2pt.FLAG_ONE_KERNEL_BOUNCE)
            assert sa.segments == sb.segments and same(a, b), (per_warp, loop_rays, depth)
    monkeypatch.delenv("B2PT_TAIL_RAYS_PER_WARP")
    monkeypatch.delenv("B2PT_TAIL_LOOP_RAYS")
    monkeypatch.delenv("B2PT_BATCH_PATHS")
    gpu_ctx.set_camera(b2pt.Camera(64, 48))
    for depth in (1, 2, 5, 30):
        a, sa = render(gpu_ctx, 5, depth, b2pt.FLAG_REFERENCE_STREAM)
        b, sb = render(gpu_ctx, 5, depth, b2pt.FLAG_REFERENCE_STREAM | b2pt.FLAG_ONE_KERNEL_BOUNCE)
        assert sa.segments == sb.segments and same(a, b), depth


def test_one_kernel_pipeline_views_and_memory_budget(gpu_ctx, b2pt):
    c = 278 / 555.0
    views = np.array([[c + 2.2 * np.cos(t), c, c + 2.2 * np.sin(t), c, c, c, 0, 1, 0, 40.0]
                      for t in np.linspace(0.2, 6.0, 9)], np.float32)
    a = gpu_ctx.render_views(views, 64, 64, 10, 5)
    b = gpu_ctx.render_views(views, 64, 64, 10, 5, flags=b2pt.FLAG_ONE_KERNEL_BOUNCE)
    assert same(a, b)
    # a memory budget only changes the batch split
    gpu_ctx.set_camera(b2pt.Camera(512, 512))
    want, s0 = render(gpu_ctx, 64, 8, 0)
    gpu_ctx.set_memory_budget(3 << 30)  # 3 GiB for the records in flight: 512*512*64 paths x 272 B = 4.6 GB do not fit
    try:
        got, s1 = render(gpu_ctx, 64, 8, 0)
    finally:
        gpu_ctx.set_memory_budget(0)
    assert s1.batches > s0.batches and same(got, want)
    with pytest.raises(b2pt.B2ptError):
        gpu_ctx.set_memory_budget(-1)
