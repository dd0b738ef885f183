"""ctypes binding of oracle/_ref/libb2pt_refharness.so -- the reference's OWN header-only worklets, compiled from
/root/reference against the minimal VTK-m stand-in in oracle/vtkm_min/ (see oracle/ref_harness.cxx).

TEST INFRASTRUCTURE ONLY: used by tests/ to pin the C oracle against the reference's code, and by
tests/golden/make_golden.py to generate fixtures.  /root/reference exists only in the build container; on the GPU
box the prebuilt library in oracle/_ref/ (git-ignored, shipped) is loaded if present.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from . import oracle as O

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libb2pt_refharness.so")
REF_ROOT = os.environ.get("B2PT_REF_ROOT", "/root/reference")


def build(force=False):
    """Compile the harness with the committed Makefile when the reference sources are present.
    Returns the library path, or None when neither the sources nor a prebuilt library exist."""
    have_src = os.path.isdir(os.path.join(REF_ROOT, "pathtracing"))
    if have_src:
        O.build()
        args = ["make", "-C", _HERE, "-s", "ref", "reftime", "refdirect", "REF_ROOT=" + REF_ROOT]
        if force:
            args.insert(1, "-B")
        subprocess.check_call(args)
    return _LIB_PATH if os.path.exists(_LIB_PATH) else None


def available():
    return build() is not None


_lib = None
_TIMING_LIB_PATH = os.path.join(_HERE, "_ref", "libb2pt_refharness_o3.so")
_tlib = None


def timing_lib():
    """The -O3 -march=x86-64-v3 build of the same harness (bench.py's CPU arm); None when it was not built."""
    global _tlib
    if _tlib is None and build() is not None and os.path.exists(_TIMING_LIB_PATH):
        O.lib()
        L = C.CDLL(_TIMING_LIB_PATH)
        L.b2ref_render.restype = C.c_int
        L.b2ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64),
                                   C.c_void_p, C.c_void_p]
        _tlib = L
    return _tlib


def lib():
    global _lib
    if _lib is None:
        path = build()
        if path is None:
            raise RuntimeError("reference harness unavailable: no /root/reference and no prebuilt oracle/_ref")
        O.lib()  # loads libb2pt_oracle.so first (orc_raygen)
        L = C.CDLL(path)
        L.b2ref_wang32.restype = C.c_uint32
        L.b2ref_wang32.argtypes = [C.POINTER(C.c_uint32)]
        L.b2ref_randf.restype = C.c_float
        L.b2ref_randf.argtypes = [C.POINTER(C.c_uint32)]
        L.b2ref_quad_hit.restype = C.c_int
        L.b2ref_quad_hit.argtypes = [C.c_void_p] * 9
        L.b2ref_sphere_hit.restype = C.c_int
        L.b2ref_sphere_hit.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_void_p, C.c_float, C.c_void_p]
        L.b2ref_raygen.restype = None
        L.b2ref_raygen.argtypes = [C.c_void_p, C.c_int64, C.POINTER(C.c_uint32), C.c_void_p]
        L.b2ref_cornell_scene.restype = C.c_int
        L.b2ref_cornell_scene.argtypes = [C.c_void_p] * 12 + [C.c_int64]
        L.b2ref_render.restype = C.c_int
        L.b2ref_render.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_int64),
                                   C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def wang_chain(seed, n):
    s = C.c_uint32(seed)
    return [int(lib().b2ref_wang32(C.byref(s))) for _ in range(n)]


def randf_chain(seed, n):
    s = C.c_uint32(seed)
    return [float(lib().b2ref_randf(C.byref(s))) for _ in range(n)]


def raygen(cam, idx, seed):
    """Camera::RayGen of the reference for pixel idx; returns (dir[3], advanced seed)."""
    s = C.c_uint32(seed)
    d = np.zeros(3, np.float32)
    cs = cam.c_struct()
    lib().b2ref_raygen(C.byref(cs), idx, C.byref(s), _p(d))
    return d, int(s.value)


def quad_hit(o, d, v00, v10, v11, v01):
    a = [np.ascontiguousarray(x, np.float32) for x in (o, d, v00, v10, v11, v01)]
    u, v, t = (C.c_float(0) for _ in range(3))
    h = lib().b2ref_quad_hit(*[_p(x) for x in a], C.byref(u), C.byref(v), C.byref(t))
    return bool(h), u.value, v.value, t.value


def sphere_hit(o, d, tmin, tmax, c, r):
    a = [np.ascontiguousarray(x, np.float32) for x in (o, d, c)]
    rec = np.zeros(9, np.float32)
    h = lib().b2ref_sphere_hit(_p(a[0]), _p(a[1]), tmin, tmax, _p(a[2]), r, _p(rec))
    return bool(h), rec


def cornell_scene():
    """The reference's own CornellBox::buildDataSet + extract (CornellBox.cpp:141-437), compiled from where it lies.
    Returns a dict of arrays in orc_cornell_scene's layout, trimmed to the counts the reference produced."""
    cap = 4096
    f32, i64, i32 = np.float32, np.int64, np.int32
    pts, quadIds = np.zeros((cap, 3), f32), np.zeros((cap, 5), i64)
    sphPt, sphR = np.zeros(cap, i64), np.zeros(cap, f32)
    matQ, texQ, matS, texS = (np.zeros(cap, i64) for _ in range(4))
    matType, texType, tex = np.zeros(5, i32), np.zeros(5, i32), np.zeros((4, 3), f32)
    n = np.zeros(3, i64)
    rc = lib().b2ref_cornell_scene(_p(pts), _p(quadIds), _p(sphPt), _p(sphR), _p(matQ), _p(texQ), _p(matS),
                                   _p(texS), _p(matType), _p(texType), _p(tex), _p(n), cap)
    if rc != 0:
        raise RuntimeError("b2ref_cornell_scene failed: %d" % rc)
    npt, nq, ns = (int(v) for v in n)
    return dict(pts=pts[:npt], quadIds=quadIds[:nq], sphPt=sphPt[:ns], sphR=sphR[:ns], matIdxQ=matQ[:nq],
                texIdxQ=texQ[:nq], matIdxS=matS[:ns], texIdxS=texS[:ns], matType=matType, texType=texType, tex=tex)


_DIRECT_LIB_PATH = os.path.join(_HERE, "_ref", "libb2pt_refdirect.so")
_dlib = None


def direct_lib():
    """The reference's Shade worklets (RayTracerNormals.cxx / RayTracerAlbedo.cxx) and Camera::PerspectiveRayGen,
    lifted and compiled by oracle/ref_direct.cxx."""
    global _dlib
    if _dlib is None:
        build()
        L = C.CDLL(_DIRECT_LIB_PATH)
        L.b2ref_direct_shade.restype = None
        L.b2ref_direct_shade.argtypes = [C.c_int] + [C.c_void_p] * 6
        L.b2ref_raygen_corner.restype = None
        L.b2ref_raygen_corner.argtypes = [C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        _dlib = L
    return _dlib


def direct_shade(which, n, p, light_pos, cam_pos, look_at):
    """Shade::operator() of RayTracerNormals (which=0) / RayTracerAlbedo (which=1) for one hit; returns rgba[4]."""
    a = [np.ascontiguousarray(x, np.float32) for x in (n, p, light_pos, cam_pos, look_at)]
    out = np.zeros(4, np.float32)
    direct_lib().b2ref_direct_shade(which, *[_p(x) for x in a], _p(out))
    return out


def raygen_corner(W, H, fov, look, up, idx):
    a = [np.ascontiguousarray(x, np.float32) for x in (look, up)]
    d = np.zeros(3, np.float32)
    direct_lib().b2ref_raygen_corner(W, H, fov, _p(a[0]), _p(a[1]), idx, _p(d))
    return d


def render_timed(scene, cam, spp, max_depth):
    """render() through the timing build when it exists; returns (rgba_sum, segments, build description)."""
    L = timing_lib()
    if L is None:
        rgba, seg, _, _ = render(scene, cam, spp, max_depth)
        return rgba, seg, "-O2 -ffp-contract=off (parity build)"
    n = cam.W * cam.H
    rgba, t0, hit0 = np.zeros((n, 4), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8)
    seg = C.c_int64(0)
    ss, cs = scene.c_struct(), cam.c_struct()
    L.b2ref_set_tree_variant(0)
    rc = L.b2ref_render(C.byref(ss), C.byref(cs), spp, max_depth, _p(rgba), C.byref(seg), _p(t0), _p(hit0))
    if rc != 0:
        raise RuntimeError("b2ref_render failed: %d" % rc)
    return rgba, int(seg.value), "-O3 -march=x86-64-v3 -ffp-contract=off (timing build)"


def render(scene, cam, spp, max_depth, tree_variant=0):
    """The reference's worklets through the stage order of MapperPathTracer.cxx:276-351.
    Returns (rgba_sum[N,4], segments, t0[N] closest distance of sample 0 / depth 0, hit0[N])."""
    n = cam.W * cam.H
    rgba = np.zeros((n, 4), np.float32)
    t0 = np.zeros(n, np.float32)
    hit0 = np.zeros(n, np.uint8)
    seg = C.c_int64(0)
    ss, cs = scene.c_struct(), cam.c_struct()
    lib().b2ref_set_tree_variant(int(tree_variant))
    rc = lib().b2ref_render(C.byref(ss), C.byref(cs), spp, max_depth, _p(rgba), C.byref(seg), _p(t0), _p(hit0))
    if rc != 0:
        raise RuntimeError("b2ref_render failed: %d" % rc)
    return rgba, int(seg.value), t0, hit0
