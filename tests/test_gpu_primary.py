"""Primary-ray specialisation (k_primary_prep + closest_small_masked): per-tile candidate masks and per-view quad
constants must never change a result.  Every render here is compared bit for bit with the generic per-ray candidate
filter (B2PT_FLAG_NO_PRIMARY_MASKS), which the other parity tests pin to the oracle; the depth-1 renders make every
path a primary ray, so a quad missing from a tile's mask shows up as a different image."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def render(ctx, cam, spp, depth, flags=0):
    ctx.set_camera(cam)
    ctx.render(spp, depth, flags)
    img = ctx.read_color().copy()
    return img, ctx.stats()


@pytest.mark.parametrize("W,H,spp,depth", [(256, 256, 64, 1), (256, 256, 16, 8), (48, 40, 200, 1), (96, 32, 64, 3),
                                           (1024, 1024, 4, 2), (8, 4, 500, 2)])
def test_masked_primary_rays_equal_generic_filter(gpu_ctx, b2pt, W, H, spp, depth):
    cam = b2pt.Camera(W, H)
    a, sa = render(gpu_ctx, cam, spp, depth)
    b, sb = render(gpu_ctx, cam, spp, depth, b2pt.FLAG_NO_PRIMARY_MASKS)
    assert sa.segments == sb.segments
    assert np.array_equal(bits(a), bits(b))


def test_masked_primary_rays_from_other_viewpoints(gpu_ctx, b2pt):
    """Cameras inside the box, looking along a wall (grazing planes), from behind the light, tilted and with a wide
    field of view: the superset rule must hold for frusta that cross planes and for planes behind the camera."""
    c = 278 / 555.0
    cams = [
        b2pt.Camera(128, 96, pos=[c, c, 0.5], lookAt=[c, 0.1, 0.9]),                 # inside the room
        b2pt.Camera(128, 96, pos=[0.02, c, -0.2], lookAt=[0.02, c, 1.0]),            # along the red wall
        b2pt.Camera(128, 96, pos=[c, 0.999, 0.5], lookAt=[c, 0.0, 0.5], up=(0, 0, 1)),  # from the ceiling, past the light
        b2pt.Camera(160, 64, pos=[1.9, 1.3, -1.4], lookAt=[c, c, c], up=(0.2, 1, 0.1), fov=75.0),
        b2pt.Camera(64, 64, pos=[c, c, 2.5], lookAt=[c, c, c]),                       # from behind the back wall
        b2pt.Camera(64, 64, pos=[-1.5, 0.2, 0.5], lookAt=[-0.6, 0.16, 0.52], fov=30.0),  # at the glass sphere
    ]
    for cam in cams:
        a, sa = render(gpu_ctx, cam, 96, 1)
        b, sb = render(gpu_ctx, cam, 96, 1, b2pt.FLAG_NO_PRIMARY_MASKS)
        assert sa.segments == sb.segments
        assert np.array_equal(bits(a), bits(b))
        a, _ = render(gpu_ctx, cam, 8, 6)
        b, _ = render(gpu_ctx, cam, 8, 6, b2pt.FLAG_NO_PRIMARY_MASKS)
        assert np.array_equal(bits(a), bits(b))


def test_masked_primary_rays_in_view_batches_and_reference_stream(gpu_ctx, b2pt):
    c = 278 / 555.0
    views = np.array([[c + 2.0 * np.cos(t), c + 0.3 * k, c + 2.0 * np.sin(t), c, c, c, 0, 1, 0, 40.0]
                      for k, t in enumerate(np.linspace(0.3, 5.9, 7))], np.float32)
    a = gpu_ctx.render_views(views, 64, 32, 12, 4)
    b = gpu_ctx.render_views(views, 64, 32, 12, 4, flags=b2pt.FLAG_NO_PRIMARY_MASKS)
    assert np.array_equal(bits(a), bits(b))
    cam = b2pt.Camera(64, 64)
    a, _ = render(gpu_ctx, cam, 6, 5, b2pt.FLAG_REFERENCE_STREAM)
    b, _ = render(gpu_ctx, cam, 6, 5, b2pt.FLAG_REFERENCE_STREAM | b2pt.FLAG_NO_PRIMARY_MASKS)
    assert np.array_equal(bits(a), bits(b))


def test_masked_primary_rays_on_a_small_sphere_scene(b2pt):
    """Small scenes with several spheres (kernel-parameter path): the sphere gate bits of the tile masks."""
    sc = b2pt.Scene.spheres(6)
    with b2pt.Context(0) as ctx:
        ctx.set_scene(sc)
        ctx.build_bvh()
        for cam in (b2pt.Camera(128, 64), b2pt.Camera(64, 64, pos=[0.5, 0.3, 0.5], lookAt=[0.9, 0.1, 0.2], fov=90.0)):
            a, sa = render(ctx, cam, 64, 1)
            b, sb = render(ctx, cam, 64, 1, b2pt.FLAG_NO_PRIMARY_MASKS)
            assert sa.tracePath == 0
            assert sa.segments == sb.segments
            assert np.array_equal(bits(a), bits(b))


@pytest.mark.parametrize("noise", [0.0, 6e-8, 2.5e-7, 2e-6])
def test_candidate_filter_on_perturbed_scenes_equals_no_filter(b2pt, noise):
    """The candidate filter against no filter at all (B2PT_FLAG_NO_AA: every quad behind its leaf box + the exact test) on
    the Cornell box with every vertex coordinate jittered: 0 keeps the world-frame axis groups exactly planar (tight
    distance bound, B2Frame::eaCoef = 4e-7), 6e-8 .. 2.5e-7 leaves the quads planar to the filter's tolerance but not
    exactly (general bound 4e-6), 2e-6 pushes most of them out of the filter behind their leaf boxes.  Primary hits,
    segments and the image must agree bit for bit in every case -- the light-sampled rays that leave the ceiling graze
    their own plane and are the ones the distance threshold must get right."""
    s = b2pt.Scene.cornell()
    if noise:
        rng = np.random.default_rng(7)
        s.pts = (s.pts.astype(np.float64) + rng.uniform(-noise, noise, s.pts.shape)).astype(np.float32)
    res = []
    for flags in (0, b2pt.FLAG_NO_AA):
        with b2pt.Context(0) as ctx:
            ctx.set_scene(s)
            ctx.build_bvh(flags)
            ctx.set_camera(b2pt.Camera(160, 128))
            prim, t = ctx.primary_hits()
            ctx.render(32, 50, flags)
            res.append((prim, t.view(np.uint32).copy(), ctx.read_color().view(np.uint32).copy(), ctx.stats().segments))
    a, b = res
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[3] == b[3]
    assert np.array_equal(a[2], b[2])
