"""Generates tests/golden/direct_refworklets.npz: outputs of the REFERENCE'S OWN -direct code -- the Shade worklets of
raytracing/RayTracerNormals.cxx:47-143 and RayTracerAlbedo.cxx:47-147 and Camera::PerspectiveRayGen
(pathtracing/Camera.cxx:339-423), lifted from /root/reference and compiled by oracle/ref_direct.cxx -- on seeded
random hits and on the pixels of two cameras.  tests/test_oracle_direct.py pins the oracle's restatement to them
(also live, where /root/reference is present).  usage: python tests/golden/make_golden_direct.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O, refharness as R  # noqa: E402

rng = np.random.default_rng(20261018)
K = 512
n = rng.normal(size=(K, 3)).astype(np.float32)
n /= np.linalg.norm(n, axis=1, keepdims=True).astype(np.float32)
p = rng.uniform(-0.2, 1.2, size=(K, 3)).astype(np.float32)
cam = rng.uniform(-2.0, 2.0, size=(K, 3)).astype(np.float32)
look = rng.uniform(0.0, 1.0, size=(K, 3)).astype(np.float32)
up = rng.normal(size=(K, 3)).astype(np.float32)
up /= np.linalg.norm(up, axis=1, keepdims=True).astype(np.float32)
up[:8] = (0, 1, 0)
light = (cam + np.float32(2) * up).astype(np.float32)  # RayTracerNormals.cxx:153-154: scale (2,2,2) * camera.GetUp()
normals = np.stack([R.direct_shade(0, n[k], p[k], light[k], cam[k], look[k]) for k in range(K)])
albedo = np.stack([R.direct_shade(1, n[k], p[k], light[k], cam[k], look[k]) for k in range(K)])
rays = {}
for name, c in (("default_64x48", O.Camera(64, 48)),
                ("moved_33x20", O.Camera(33, 20, pos=[1.4, 0.9, -1.1], lookAt=[0.4, 0.5, 0.5], up=(0.1, 1.0, 0.2), fov=55.0))):
    lk = (c.lookAt - c.pos).astype(np.float32)
    lk = (lk * np.float32(1.0 / np.sqrt(np.float32((lk[0] * lk[0] + lk[1] * lk[1]) + lk[2] * lk[2])))).astype(np.float32)
    upv = c.up.astype(np.float32)
    if not (upv[0] == 0 and upv[1] == 1 and upv[2] == 0):
        upv = (upv * np.float32(1.0 / np.sqrt(np.float32((upv[0] * upv[0] + upv[1] * upv[1]) + upv[2] * upv[2])))).astype(np.float32)
    rays[name] = np.stack([R.raygen_corner(c.W, c.H, c.fov, lk, upv, i) for i in range(c.W * c.H)])
np.savez_compressed(os.path.join(HERE, "direct_refworklets.npz"), n=n, p=p, cam=cam, look=look, up=up,
                    normals=normals, albedo=albedo, **{"rays_" + k: v for k, v in rays.items()})
print("wrote direct_refworklets.npz")
