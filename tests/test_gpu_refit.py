"""b2pt_update_spheres + b2pt_refit_bvh (SURVEY.md 8f-1 "+ refit"): the tree keeps its topology, every box is refitted
bottom-up on the device.  Closest hits do not depend on the tree, so a refitted tree must give the hits and the image
of a fresh build over the moved scene -- for both builders and for the kernel-parameter path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def moved(scene, b2pt, rng, amount):
    s = b2pt.Scene.spheres(len(scene.sphPt))
    c = s.pts[: len(s.sphPt)] + rng.uniform(-amount, amount, size=(len(s.sphPt), 3)).astype(np.float32)
    s.pts[: len(s.sphPt)] = c
    return s, c


@pytest.mark.parametrize("n,flags_name", [(6, "none"), (700, "none"), (20000, "none"), (20000, "lbvh")])
def test_refit_equals_rebuild(b2pt, n, flags_name):
    flags = b2pt.FLAG_GPU_LBVH if flags_name == "lbvh" else 0
    rng = np.random.default_rng(n)
    base = b2pt.Scene.spheres(n)
    W, H, spp, depth = 96, 54, 4, 8
    with b2pt.Context(0) as ctx:
        ctx.set_scene(base)
        ctx.build_bvh(flags)
        ctx.set_camera(b2pt.Camera(W, H))
        p0, t0 = ctx.primary_hits()
        for step, amount in enumerate((0.003, 0.02)):
            new_scene, centers = moved(base, b2pt, rng, amount)
            ctx.update_spheres(centers)
            ctx.refit_bvh()
            p1, t1 = ctx.primary_hits()
            ctx.render(spp, depth, flags)
            a, sa = ctx.read_color().copy(), ctx.stats()
            with b2pt.Context(0) as fresh:
                fresh.set_scene(new_scene)
                fresh.build_bvh(flags)
                fresh.set_camera(b2pt.Camera(W, H))
                p2, t2 = fresh.primary_hits()
                fresh.render(spp, depth, flags)
                b, sb = fresh.read_color().copy(), fresh.stats()
            assert not np.array_equal(p0, p1) or n < 10  # the spheres really moved
            assert np.array_equal(p1, p2) and np.array_equal(t1.view(np.uint32), t2.view(np.uint32)), (step, amount)
            assert sa.segments == sb.segments
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_refit_errors(b2pt):
    with b2pt.Context(0) as ctx:
        with pytest.raises(b2pt.B2ptError):
            ctx.refit_bvh()  # nothing built
        ctx.set_scene(b2pt.Scene.spheres(100))
        ctx.build_bvh(b2pt.FLAG_WIDE_BVH)
        with pytest.raises(b2pt.B2ptError):
            ctx.refit_bvh()  # binary tree only
        ctx.build_bvh(0)
        ctx.refit_bvh()
