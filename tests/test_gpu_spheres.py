"""BASELINE.json configs[3]: the synthetic many-sphere scene through the BVH traversal kernels.

There is no reference counterpart for more than one sphere (SphereExtractor.cxx:108-111), so parity is against
the oracle's brute-force closest hit (orc_closest_hit tests every primitive in index order)."""
import ctypes as C
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _oracle_scene(oracle, s):
    return oracle.Scene(s.pts, s.quadIds, s.sphPt, s.sphR, s.matIdxQ, s.texIdxQ, s.matIdxS, s.texIdxS, s.matType,
                        s.texType, s.tex, s.lightQuadIds, s.lightSphPt, s.lightSphR, s.lightables, s.refIdx)


@pytest.mark.parametrize("n,W,H,spp,depth", [(300, 64, 36, 4, 6), (5000, 96, 54, 2, 8)])
def test_sphere_scene_matches_oracle(b2pt, oracle, n, W, H, spp, depth):
    s = b2pt.Scene.spheres(n)
    osc, ocam = _oracle_scene(oracle, s), oracle.Camera(W, H)
    with b2pt.Context(0) as ctx:
        ctx.set_scene(s)
        ctx.build_bvh()
        ctx.set_camera(b2pt.Camera(W, H))
        prim, t = ctx.primary_hits()
        ctx.render(spp, depth, 0)
        g, st = ctx.read_color(), ctx.stats()
    assert st.tracePath == 1 and st.bvhNodes > 1  # BVH kernels, not the small-scene path
    oprim, ot = oracle.primary_hits(osc, ocam)
    assert np.array_equal(prim, oprim)
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    o, ost = oracle.render(osc, ocam, spp, depth, mode=oracle.MODE_FORWARD_FAST)
    assert st.segments == ost.segments  # identical trajectories
    ok = ~(np.isnan(o[:, :3]) | np.isnan(g[:, :3]))
    rel = np.abs(g[:, :3][ok] - o[:, :3][ok]) / np.maximum(np.abs(o[:, :3][ok]), 1e-3 * spp)
    assert (rel < 1e-4).mean() > 0.9995


def test_million_sphere_scene_properties(b2pt, oracle):
    """Full-size configs[3]: 1,000,000 spheres, 1920x1080.  Size-independent checks: sampled primary hits equal the
    oracle's brute force over all primitives bit for bit, renders are deterministic and additive over sample ranges."""
    n, W, H = 1_000_000, 1920, 1080
    s = b2pt.Scene.spheres(n)
    with b2pt.Context(0) as ctx:
        ctx.set_scene(s)
        t0 = time.time()
        ctx.build_bvh()
        build_s = time.time() - t0
        ctx.set_camera(b2pt.Camera(W, H))
        prim, t = ctx.primary_hits()
        ctx.render(4, 50, 0)
        a, st = ctx.read_color(), ctx.stats()
        ctx.clear_color()
        ctx.render_range(0, 1, 50, 0)
        ctx.render_range(1, 3, 50, 0)
        b = ctx.read_color()
    assert st.tracePath == 1 and st.tracedSpheres == n
    assert np.array_equal(a, b, equal_nan=True)
    assert st.segments > st.paths
    assert build_s < 120
    # brute-force check of a pixel sample (every primitive tested, index order)
    osc, ocam = _oracle_scene(oracle, s), oracle.Camera(W, H)
    rng = np.random.default_rng(3)
    pix = rng.choice(W * H, 192, replace=False)
    for p in pix:
        d, _ = oracle.raygen(ocam, int(p), int(p))
        oprim, rec, _ = oracle.closest_hit(osc, ocam.pos, d)
        assert oprim == prim[p], (p, oprim, prim[p])
        if oprim >= 0:
            assert np.float32(rec[2]).view(np.uint32) == t[p].view(np.uint32)
    assert (prim >= 2).mean() > 0.2  # spheres are actually visible


@pytest.mark.parametrize("n,W,H,spp,depth", [(1, 32, 18, 2, 4), (3, 32, 18, 2, 4), (700, 64, 36, 4, 6), (20000, 96, 54, 2, 8)])
def test_gpu_lbvh_builder_matches_oracle(b2pt, oracle, n, W, H, spp, depth):
    """B2PT_FLAG_GPU_LBVH: Morton order + Karras radix tree + atomic bottom-up fit on the device.  A different tree,
    the same closest hits: primary ids / t bit-identical to brute force, trajectories identical."""
    s = b2pt.Scene.spheres(n)
    osc, ocam = _oracle_scene(oracle, s), oracle.Camera(W, H)
    with b2pt.Context(0) as ctx:
        ctx.set_scene(s)
        ctx.build_bvh(b2pt.FLAG_GPU_LBVH | b2pt.FLAG_FORCE_BVH)
        ctx.set_camera(b2pt.Camera(W, H))
        prim, t = ctx.primary_hits()
        ctx.render(spp, depth, b2pt.FLAG_GPU_LBVH | b2pt.FLAG_FORCE_BVH)
        g, st = ctx.read_color(), ctx.stats()
    assert st.tracePath == 1 and st.bvhNodes == 2 * (n + 2)
    oprim, ot = oracle.primary_hits(osc, ocam)
    assert np.array_equal(prim, oprim)
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    o, ost = oracle.render(osc, ocam, spp, depth, mode=oracle.MODE_FORWARD_FAST)
    # Among tens of thousands of tiny spheres a 1-ulp difference between CUDA's sincosf and glibc's sinf/cosf in a
    # sampled direction occasionally grazes a different sphere: the paths agree except for a handful (the same
    # handful for the host-SAH tree, see below), so the segment count is compared with a tolerance at this size.
    assert abs(st.segments - ost.segments) <= (0 if n <= 700 else 2e-3 * ost.segments)
    ok = ~(np.isnan(o[:, :3]) | np.isnan(g[:, :3]))
    rel = np.abs(g[:, :3][ok] - o[:, :3][ok]) / np.maximum(np.abs(o[:, :3][ok]), 1e-3 * spp)
    assert (rel < 1e-4).mean() > (0.9995 if n <= 700 else 0.99)
    # the host-SAH tree traces exactly the same paths
    with b2pt.Context(0) as ctx:
        ctx.set_scene(s)
        ctx.build_bvh(b2pt.FLAG_FORCE_BVH)
        ctx.set_camera(b2pt.Camera(W, H))
        ctx.render(spp, depth, b2pt.FLAG_FORCE_BVH)
        g2, st2 = ctx.read_color(), ctx.stats()
    assert st2.segments == st.segments and np.array_equal(g, g2, equal_nan=True)


def test_gpu_lbvh_equals_host_sah_on_the_million_sphere_scene(b2pt):
    """Both builders, 1M spheres: identical primary hits and identical images.  Build times are printed, not asserted
    (wall-clock comparisons flake on a shared box; scripts/time_build.py measures them)."""
    n, W, H = 1_000_000, 960, 540
    s = b2pt.Scene.spheres(n)
    with b2pt.Context(0) as ctx:
        ctx.set_scene(s)
        ctx.set_camera(b2pt.Camera(W, H))
        t0 = time.time()
        ctx.build_bvh(0)
        ctx.synchronize()
        host_s = time.time() - t0
        p0, t0_ = ctx.primary_hits()
        ctx.render(2, 50, 0)
        a, sa = ctx.read_color(), ctx.stats()
        t1 = time.time()
        ctx.build_bvh(b2pt.FLAG_GPU_LBVH)
        ctx.synchronize()
        gpu_s = time.time() - t1
        p1, t1_ = ctx.primary_hits()
        ctx.render(2, 50, b2pt.FLAG_GPU_LBVH)
        b, sb = ctx.read_color(), ctx.stats()
    # Tree independence holds except for near-coincident grazing hits: the reference's sphere formula cancels badly
    # for small, far spheres and can report a hit whose t lies BEFORE the ray enters the sphere's own box; whether a
    # traversal that prunes by the closest distance still reaches that box depends on the order of the visits
    # (measured: 2 of 518 400 primary rays differ between the two trees).
    same = (p0 == p1) & (t0_.view(np.uint32) == t1_.view(np.uint32))
    assert (~same).sum() <= 1e-5 * same.size, int((~same).sum())
    assert abs(sa.segments - sb.segments) <= 1e-5 * sa.segments
    eq = ((a == b) | (np.isnan(a) & np.isnan(b))).all(1)
    assert eq.mean() > 0.9999
    assert sb.bvhNodes == 2 * (n + 2)
    print("host SAH build %.3f s, device LBVH build %.3f s" % (host_s, gpu_s))


@pytest.mark.parametrize("n,W,H,spp,depth", [(9, 64, 36, 8, 6), (300, 64, 36, 8, 6), (20000, 192, 108, 8, 12)])
def test_wide_tree_and_ray_sort_change_nothing(b2pt, n, W, H, spp, depth, monkeypatch):
    """The 8-wide compressed tree against the binary tree it was collapsed from, the spatially sorted ray order
    against queue order, and the split bounce (k_bvh_hits + k_resolve_hits) against the single k_trace: different node
    visits, different warps, different kernels -- the same closest hits, the same paths, the same image, bit for bit
    (several batches in flight so that the tail modes take part)."""
    s = b2pt.Scene.spheres(n)
    monkeypatch.setenv("B2PT_BATCH_PATHS", str(W * H * 2))
    monkeypatch.setenv("B2PT_VALIDATE_BVH", "1")
    ref = None
    F = b2pt.FLAG_FORCE_BVH
    for flags in (F, F | b2pt.FLAG_WIDE_BVH, F | b2pt.FLAG_NO_RAY_SORT, F | b2pt.FLAG_WIDE_BVH | b2pt.FLAG_NO_RAY_SORT,
                  F | b2pt.FLAG_SPLIT_TRACE, F | b2pt.FLAG_NO_TAIL, F | b2pt.FLAG_NO_TAIL | b2pt.FLAG_SPLIT_TRACE,
                  F | b2pt.FLAG_GPU_LBVH, F | b2pt.FLAG_GPU_LBVH | b2pt.FLAG_WIDE_BVH,
                  F | b2pt.FLAG_GPU_LBVH | b2pt.FLAG_SPLIT_TRACE):
        with b2pt.Context(0) as ctx:
            ctx.set_scene(s)
            ctx.build_bvh(flags)
            ctx.set_camera(b2pt.Camera(W, H))
            prim, t = ctx.primary_hits()
            ctx.render(spp, depth, flags)
            img, st = ctx.read_color().copy(), ctx.stats()
        assert st.tracePath == 1 and st.batches == 4
        cur = (prim, t.view(np.uint32), img.view(np.uint32), st.segments)
        if ref is None:
            ref = cur
        else:
            assert np.array_equal(ref[0], cur[0]) and np.array_equal(ref[1], cur[1]), flags
            assert ref[3] == cur[3], flags
            assert np.array_equal(ref[2], cur[2]), flags
