"""-direct G-buffers (main.cc:402-422): the oracle's restatement of the per-pixel rules against the
reference's own Shade worklets (RayTracerNormals.cxx:47-143, RayTracerAlbedo.cxx:47-147) and pixel-ray generator
(Camera::PerspectiveRayGen, pathtracing/Camera.cxx:339-423) -- through the committed fixture everywhere, and live where
/root/reference is present (oracle/ref_direct.cxx compiles the lifted classes)."""
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _light_up(up):
    return up  # the fixture's `up` vectors are already unit length (or the default (0,1,0))


def test_shade_rules_match_reference_fixture(oracle):
    z = np.load(os.path.join(HERE, "golden", "direct_refworklets.npz"))
    K = len(z["n"])
    for k in range(K):
        nrm, alb = oracle.direct_shade(z["n"][k], z["p"][k], z["cam"][k], z["look"][k], _light_up(z["up"][k]))
        assert np.array_equal(bits(nrm), bits(z["normals"][k])), k
        assert np.array_equal(bits(alb), bits(z["albedo"][k])), k  # includes inf / NaN where cosTheta clamps to 0


def test_pixel_rays_match_reference_fixture(oracle):
    z = np.load(os.path.join(HERE, "golden", "direct_refworklets.npz"))
    cams = {"default_64x48": oracle.Camera(64, 48),
            "moved_33x20": oracle.Camera(33, 20, pos=[1.4, 0.9, -1.1], lookAt=[0.4, 0.5, 0.5], up=(0.1, 1.0, 0.2), fov=55.0)}
    for name, cam in cams.items():
        want = z["rays_" + name]
        got = np.stack([oracle.raygen_corner(cam, i) for i in range(cam.W * cam.H)])
        assert np.array_equal(bits(got), bits(want)), name


def test_live_shade_rules_match_reference_classes(oracle):
    from oracle import refharness as R
    if not R.available():
        pytest.skip("reference sources not present")
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = rng.normal(size=3).astype(np.float32)
        n /= np.float32(np.linalg.norm(n))
        p, cam, look = (rng.uniform(-1, 2, 3).astype(np.float32) for _ in range(3))
        up = rng.normal(size=3).astype(np.float32)
        up /= np.float32(np.linalg.norm(up))
        light = (cam + np.float32(2) * up).astype(np.float32)
        nrm, alb = oracle.direct_shade(n, p, cam, look, up)
        assert np.array_equal(bits(nrm), bits(R.direct_shade(0, n, p, light, cam, look)))
        assert np.array_equal(bits(alb), bits(R.direct_shade(1, n, p, light, cam, look)))


def test_direct_buffers_of_the_cornell_box(oracle):
    """Whole-image sanity of orc_direct: quads only (the glass sphere is not part of the MapperQuad scene), unit
    normals opposing the ray, depth = hit distance, misses untouched."""
    sc, cam = oracle.cornell_scene(), oracle.Camera(96, 64)
    normals, albedo, depth, prim = oracle.direct(sc, cam)
    hit = prim >= 0
    assert 0.5 < hit.mean() < 1.0 and prim.max() < 22
    assert np.allclose(np.linalg.norm(normals[hit, :3], axis=1), 1.0, atol=1e-5) and (normals[hit, 3] == 1).all()
    assert not normals[~hit].any() and not albedo[~hit].any() and not depth[~hit].any()
    assert (depth[hit] > 0.5).all() and (albedo[hit, 3] == 1).all()
    d = np.stack([oracle.raygen_corner(cam, int(i)) for i in np.flatnonzero(hit)[:200]])
    assert ((normals[hit][:200, :3] * d).sum(1) <= 1e-6).all()  # flipped to oppose the ray
