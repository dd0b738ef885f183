"""-direct G-buffers on the GPU (b2pt_render_direct, k_direct) against the oracle's restatement, which
tests/test_oracle_direct.py pins bit for bit to the reference's Shade worklets and pixel-ray generator."""
import os
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("W,H,kw", [(96, 64, {}), (128, 128, {}), (33, 20, dict(pos=[1.4, 0.9, -1.1], lookAt=[0.4, 0.5, 0.5],
                                                                                  up=(0.1, 1.0, 0.2), fov=55.0)),
                                     (64, 64, dict(pos=[0.5, 0.5, 0.5], lookAt=[0.9, 0.1, 0.9], fov=70.0))])
def test_direct_buffers_match_oracle(gpu_ctx, b2pt, oracle, W, H, kw):
    gpu_ctx.set_camera(b2pt.Camera(W, H, **kw))
    normals, albedo, depth, prim = gpu_ctx.render_direct()
    on, oa, od, op = oracle.direct(oracle.cornell_scene(), oracle.Camera(W, H, **kw))
    assert np.array_equal(prim, op)
    assert np.array_equal(bits(depth), bits(od))        # the exact quad test: same t
    assert np.array_equal(bits(normals), bits(on))      # host-precomputed unit normals, flipped to oppose the ray
    assert np.array_equal(bits(albedo), bits(oa))       # same float operations in the same order (no FMA contraction)
    assert (prim < 22).all() and (prim >= 0).mean() > 0.3


def test_direct_mode_of_the_driver(b2pt, oracle, tmp_path):
    """CornellBox_b2pt -direct writes the reference's normals / albedo / depth PNM files (main.cc:402-422, save())."""
    exe = os.path.join(ROOT, "raytracingtherestofyourlife_b200", "host", "CornellBox_b2pt")
    out = subprocess.run([exe, "-direct", "-x", "48", "-y", "40"], cwd=tmp_path, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    on, oa, od, _ = oracle.direct(oracle.cornell_scene(), oracle.Camera(48, 40))

    def read(name):
        toks = open(tmp_path / name).read().split()
        assert toks[:4] == ["P3", "48", "40", "255"]
        return np.array(toks[4:], np.int64).reshape(-1, 3)

    def save_vec4(c):  # main.cc:331-341
        c = c.copy()
        c[np.isnan(c[:, :3]).any(1)] = 0
        with np.errstate(invalid="ignore", over="ignore"):
            return np.trunc(255.99 * c[:, :3].astype(np.float64))

    n_img, a_img, d_img = read("normals.pnm"), read("albedo.pnm"), read("depth.pnm")
    assert np.array_equal(n_img, save_vec4(on).astype(np.int64))
    fin = np.isfinite(save_vec4(oa)).all(1) & (np.abs(save_vec4(oa)) < 2**31 - 1).all(1)
    assert np.array_equal(a_img[fin], save_vec4(oa)[fin].astype(np.int64))
    assert np.array_equal(d_img[:, 0], np.trunc(255.99 * np.sqrt(od).astype(np.float32).astype(np.float64)).astype(np.int64))
