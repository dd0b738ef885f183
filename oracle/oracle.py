"""ctypes binding of the CPU ORACLE (oracle/b2pt_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.
Parity status: pinned to the reference's worklet code through oracle/ref_harness.cxx (see b2pt_oracle.h).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libb2pt_oracle.so")

MODE_PASSES, MODE_FUSED, MODE_FORWARD_BURN, MODE_FORWARD_FAST = 0, 1, 2, 3
FLAG_KILL_ZERO_THROUGHPUT, FLAG_NO_AABB_GATE = 1, 2


def build(force=False):
    """Compile the oracle with the committed Makefile (gcc only, no reference sources)."""
    src = [os.path.join(_HERE, f) for f in ("b2pt_oracle.c", "b2pt_oracle.h", "Makefile")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class _Scene(C.Structure):
    _fields_ = [
        ("nPts", C.c_int64), ("pts", C.c_void_p),
        ("nQuads", C.c_int64), ("quadIds", C.c_void_p),
        ("nSph", C.c_int64), ("sphPt", C.c_void_p), ("sphR", C.c_void_p),
        ("matIdxQ", C.c_void_p), ("texIdxQ", C.c_void_p), ("matIdxS", C.c_void_p), ("texIdxS", C.c_void_p),
        ("nMatType", C.c_int), ("matType", C.c_void_p),
        ("nTexType", C.c_int), ("texType", C.c_void_p),
        ("nTex", C.c_int), ("tex", C.c_void_p),
        ("nLightQuads", C.c_int64), ("lightQuadIds", C.c_void_p),
        ("nLightSph", C.c_int64), ("lightSphPt", C.c_void_p), ("lightSphR", C.c_void_p),
        ("lightables", C.c_int), ("refIdx", C.c_float),
    ]


class _Camera(C.Structure):
    _fields_ = [("pos", C.c_float * 3), ("lookAt", C.c_float * 3), ("up", C.c_float * 3),
                ("fovDeg", C.c_float), ("W", C.c_int), ("H", C.c_int)]


class Stats(C.Structure):
    _fields_ = [("paths", C.c_int64), ("segments", C.c_int64), ("rngDraws", C.c_int64),
                ("nanSamples", C.c_int64), ("zeroKilled", C.c_int64), ("aliveAtDepth", C.c_int64 * 64)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_wang32.restype = C.c_uint32
        L.orc_wang32.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_randf.restype = C.c_float
        L.orc_randf.argtypes = [C.POINTER(C.c_uint32)]
        L.orc_wang_init.restype = C.c_uint32
        L.orc_wang_init.argtypes = [C.c_uint32]
        L.orc_cornell_scene.restype = C.c_int
        L.orc_cornell_scene.argtypes = [C.c_void_p] * 11
        L.orc_camera_basis.argtypes = [C.POINTER(_Camera), C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_raygen.argtypes = [C.POINTER(_Camera), C.c_int64, C.POINTER(C.c_uint32), C.c_void_p]
        L.orc_closest_hit.restype = C.c_int64
        L.orc_closest_hit.argtypes = [C.POINTER(_Scene), C.c_void_p, C.c_void_p, C.c_float, C.c_float, C.c_int,
                                      C.c_void_p, C.c_void_p]
        L.orc_primary_hits.restype = C.c_int
        L.orc_primary_hits.argtypes = [C.POINTER(_Scene), C.POINTER(_Camera), C.c_uint32, C.c_int, C.c_void_p,
                                       C.c_void_p]
        L.orc_render.restype = C.c_int
        L.orc_render.argtypes = [C.POINTER(_Scene), C.POINTER(_Camera), C.c_int, C.c_int, C.c_int, C.c_uint32,
                                 C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(Stats)]
        L.orc_trace_path.restype = C.c_int
        L.orc_trace_path.argtypes = [C.POINTER(_Scene), C.POINTER(_Camera), C.c_int64, C.c_uint32, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p]
        L.orc_normalize.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
        L.orc_quad_pdf_value.restype = C.c_float
        L.orc_quad_pdf_value.argtypes = [C.c_void_p] * 6
        L.orc_sphere_pdf_value.restype = C.c_float
        L.orc_sphere_pdf_value.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_float]
        L.orc_quad_hit.restype = C.c_int
        L.orc_quad_hit.argtypes = [C.c_void_p] * 9
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def wang_chain(seed, n):
    """n successive getWang32 outputs starting from state `seed` (wangXor.h:30-38)."""
    s = C.c_uint32(seed)
    return [int(lib().orc_wang32(C.byref(s))) for _ in range(n)]


def randf_chain(seed, n):
    s = C.c_uint32(seed)
    return [float(lib().orc_randf(C.byref(s))) for _ in range(n)]


def wang_init(x):
    return int(lib().orc_wang_init(x))


class Scene:
    """Plain-array scene, the inputs of MapperPathTracer::RenderCells."""

    def __init__(self, pts, quadIds, sphPt, sphR, matIdxQ, texIdxQ, matIdxS, texIdxS, matType, texType, tex,
                 lightQuadIds, lightSphPt, lightSphR, lightables=2, refIdx=1.5):
        f32, i64, i32 = np.float32, np.int64, np.int32
        self.pts = np.ascontiguousarray(pts, f32).reshape(-1, 3)
        self.quadIds = np.ascontiguousarray(quadIds, i64).reshape(-1, 5)
        self.sphPt = np.ascontiguousarray(sphPt, i64).reshape(-1)
        self.sphR = np.ascontiguousarray(sphR, f32).reshape(-1)
        self.matIdxQ = np.ascontiguousarray(matIdxQ, i64).reshape(-1)
        self.texIdxQ = np.ascontiguousarray(texIdxQ, i64).reshape(-1)
        self.matIdxS = np.ascontiguousarray(matIdxS, i64).reshape(-1)
        self.texIdxS = np.ascontiguousarray(texIdxS, i64).reshape(-1)
        self.matType = np.ascontiguousarray(matType, i32).reshape(-1)
        self.texType = np.ascontiguousarray(texType, i32).reshape(-1)
        self.tex = np.ascontiguousarray(tex, f32).reshape(-1, 3)
        self.lightQuadIds = np.ascontiguousarray(lightQuadIds, i64).reshape(-1, 5)
        self.lightSphPt = np.ascontiguousarray(lightSphPt, i64).reshape(-1)
        self.lightSphR = np.ascontiguousarray(lightSphR, f32).reshape(-1)
        self.lightables = int(lightables)
        self.refIdx = float(refIdx)

    def c_struct(self):
        s = _Scene()
        s.nPts, s.pts = len(self.pts), _p(self.pts)
        s.nQuads, s.quadIds = len(self.quadIds), _p(self.quadIds)
        s.nSph, s.sphPt, s.sphR = len(self.sphPt), _p(self.sphPt), _p(self.sphR)
        s.matIdxQ, s.texIdxQ, s.matIdxS, s.texIdxS = _p(self.matIdxQ), _p(self.texIdxQ), _p(self.matIdxS), _p(
            self.texIdxS)
        s.nMatType, s.matType = len(self.matType), _p(self.matType)
        s.nTexType, s.texType = len(self.texType), _p(self.texType)
        s.nTex, s.tex = len(self.tex), _p(self.tex)
        s.nLightQuads, s.lightQuadIds = len(self.lightQuadIds), _p(self.lightQuadIds)
        s.nLightSph, s.lightSphPt, s.lightSphR = len(self.lightSphPt), _p(self.lightSphPt), _p(self.lightSphR)
        s.lightables, s.refIdx = self.lightables, self.refIdx
        return s


def cornell_scene():
    """CornellBox.cpp:141-418 plus the light lists of MapperPathTracer.cxx:141-148."""
    f32, i64, i32 = np.float32, np.int64, np.int32
    pts = np.zeros((89, 3), f32)
    quadIds = np.zeros((22, 5), i64)
    sphPt, sphR = np.zeros(1, i64), np.zeros(1, f32)
    matQ, texQ = np.zeros(22, i64), np.zeros(22, i64)
    matS, texS = np.zeros(1, i64), np.zeros(1, i64)
    matType, texType = np.zeros(5, i32), np.zeros(5, i32)
    tex = np.zeros((4, 3), f32)
    rc = lib().orc_cornell_scene(_p(pts), _p(quadIds), _p(sphPt), _p(sphR), _p(matQ), _p(texQ), _p(matS), _p(texS),
                                 _p(matType), _p(texType), _p(tex))
    assert rc == 0
    return Scene(pts, quadIds, sphPt, sphR, matQ, texQ, matS, texS, matType, texType, tex,
                 lightQuadIds=[[0, 8, 9, 10, 11]], lightSphPt=[48], lightSphR=[sphR[0]])


class Camera:
    """main.cc:616-622 defaults."""

    def __init__(self, W, H, pos=None, lookAt=None, up=(0, 1, 0), fov=40.0):
        f = np.float32
        self.pos = np.array(pos if pos is not None else [278 / 555.0, 278 / 555.0, -800 / 555.0], f)
        self.lookAt = np.array(lookAt if lookAt is not None else [278 / 555.0, 278 / 555.0, 278 / 555.0], f)
        self.up = np.array(up, f)
        self.fov, self.W, self.H = float(fov), int(W), int(H)

    def c_struct(self):
        c = _Camera()
        for k in range(3):
            c.pos[k], c.lookAt[k], c.up[k] = float(self.pos[k]), float(self.lookAt[k]), float(self.up[k])
        c.fovDeg, c.W, c.H = self.fov, self.W, self.H
        return c

    def basis(self):
        a, b, c = (np.zeros(3, np.float32) for _ in range(3))
        cs = self.c_struct()
        lib().orc_camera_basis(C.byref(cs), _p(a), _p(b), _p(c))
        return a, b, c


def raygen(cam, idx, seed):
    s = C.c_uint32(seed)
    d = np.zeros(3, np.float32)
    cs = cam.c_struct()
    lib().orc_raygen(C.byref(cs), idx, C.byref(s), _p(d))
    return d, int(s.value)


def closest_hit(scene, o, d, tmin=0.001, tmax=np.finfo(np.float32).max, flags=0):
    o = np.ascontiguousarray(o, np.float32)
    d = np.ascontiguousarray(d, np.float32)
    rec = np.zeros(9, np.float32)
    hid = np.zeros(2, np.int32)
    ss = scene.c_struct()
    prim = lib().orc_closest_hit(C.byref(ss), _p(o), _p(d), tmin, tmax, flags, _p(rec), _p(hid))
    return int(prim), rec, hid


def primary_hits(scene, cam, seed_offset=0, flags=0):
    n = cam.W * cam.H
    prim = np.zeros(n, np.int32)
    t = np.zeros(n, np.float32)
    ss, cs = scene.c_struct(), cam.c_struct()
    rc = lib().orc_primary_hits(C.byref(ss), C.byref(cs), seed_offset, flags, _p(prim), _p(t))
    assert rc == 0
    return prim, t


def render(scene, cam, spp, max_depth, mode=MODE_FORWARD_FAST, sample_begin=0, seed_offset=0, flags=0, threads=0):
    """Returns (rgba_sum[H*W,4] float32, Stats)."""
    n = cam.W * cam.H
    rgba = np.zeros((n, 4), np.float32)
    st = Stats()
    ss, cs = scene.c_struct(), cam.c_struct()
    rc = lib().orc_render(C.byref(ss), C.byref(cs), spp, sample_begin, max_depth, seed_offset, mode, flags, threads,
                          _p(rgba), C.byref(st))
    if rc != 0:
        raise RuntimeError("orc_render failed: %d" % rc)
    return rgba, st


def direct(scene, cam):
    """-direct G-buffers (orc_direct): (normals[N,4], albedo[N,4], depth[N], prim[N])."""
    n = cam.W * cam.H
    normals, albedo = np.zeros((n, 4), np.float32), np.zeros((n, 4), np.float32)
    depth, prim = np.zeros(n, np.float32), np.zeros(n, np.int32)
    ss, cs = scene.c_struct(), cam.c_struct()
    L = lib()
    L.orc_direct.restype = C.c_int
    rc = L.orc_direct(C.byref(ss), C.byref(cs), _p(normals), _p(albedo), _p(depth), _p(prim))
    assert rc == 0
    return normals, albedo, depth, prim


def direct_shade(n, p, cam_pos, look_at, up_n):
    """The two Shade rules for one hit (orc_direct_shade): (normals4, albedo4)."""
    a = [np.ascontiguousarray(x, np.float32) for x in (n, p, cam_pos, look_at, up_n)]
    o0, o1 = np.zeros(4, np.float32), np.zeros(4, np.float32)
    lib().orc_direct_shade(*[_p(x) for x in a], _p(o0), _p(o1))
    return o0, o1


def raygen_corner(cam, idx):
    d = np.zeros(3, np.float32)
    cs = cam.c_struct()
    lib().orc_raygen_corner(C.byref(cs), C.c_int64(idx), _p(d))
    return d


def normalize(rgba_sum, spp):
    out = np.zeros_like(rgba_sum)
    lib().orc_normalize(_p(np.ascontiguousarray(rgba_sum)), rgba_sum.shape[0], spp, _p(out))
    return out


def trace_path(scene, cam, pixel, rng_state, max_depth, flags=0):
    """One forward path sample with a per-depth log [max_depth, 20]; returns (L[3], log, segments)."""
    L = np.zeros(3, np.float32)
    log = np.zeros((max_depth, 20), np.float32)
    ss, cs = scene.c_struct(), cam.c_struct()
    n = lib().orc_trace_path(C.byref(ss), C.byref(cs), pixel, rng_state & 0xFFFFFFFF, max_depth, flags, _p(L), _p(log))
    return L, log, n
