"""The C++ drop-in facade (raytracingtherestofyourlife_b200/host): MapperPathTracer / PathTracer / Camera /
ChannelBuffer / Ray with the reference's signatures over the C-ABI.  The C++ checks live in host/test_facade.cc;
this module builds and runs them and compares GPU outputs with the oracle."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "raytracingtherestofyourlife_b200", "host")


@pytest.fixture(scope="module")
def host_build(b2pt):
    import runpy
    runpy.run_path(os.path.join(HOST, "build_host.py"), run_name="__test_build__")["build_host"]()
    return HOST


def test_facade_host_behaviour(host_build):
    """Containers, validation and error messages: no GPU needed."""
    r = subprocess.run([os.path.join(host_build, "test_facade"), "--cpu"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all facade checks passed" in r.stdout


def test_facade_headers_keep_the_reference_surface():
    """Spot-check that the class surface named in SURVEY.md 8b is declared."""
    hdr = open(os.path.join(HOST, "MapperPathTracer.h")).read()
    for sig in ("class MapperPathTracer : public Mapper", "void SetCanvas(vtkm::rendering::Canvas* canvas) override",
                "void RenderCells(const vtkm::cont::DynamicCellSet& cellset", "void buildBVH(", "void intersect(",
                "void applyMaterials(", "void applyPDFs(", "void generateRays(", "extract(const vtkm::cont::DynamicCellSet",
                "const int depthcount, samplecount;", "whichPDF", "srecs;", "hrecs;", "hids;"):
        assert sig in hdr, sig
    cam = open(os.path.join(HOST, "pathtracing", "Camera.h")).read()
    for sig in ("void SetParameters(const vtkm::rendering::Camera& camera, vtkm::rendering::CanvasRayTracer& canvas)",
                "vtkm::cont::ArrayHandle<unsigned int> seeds;", "void CreateRays(vtkm::rendering::raytracing::Ray<vtkm::Float32>&",
                "void CreateRays(vtkm::rendering::raytracing::Ray<vtkm::Float64>&", "GetSubsetWidth", "ResetIsViewDirty",
                "bool operator==(const Camera& other) const"):
        assert sig in cam, sig


def test_facade_calls_only_the_c_abi():
    """The facade may include nothing of the CUDA implementation: only include/b2pt.h."""
    for d, _, files in os.walk(HOST):
        if "vtkm_shim" in d:
            continue
        for f in files:
            if f.endswith((".h", ".cxx", ".cpp", ".cc")):
                src = open(os.path.join(d, f)).read()
                assert "cuda_runtime" not in src and "csrc/" not in src and "<<<" not in src, f


@pytest.mark.gpu
def test_facade_render_matches_oracle(host_build, tmp_path, oracle):
    prefix = str(tmp_path / "facade")
    r = subprocess.run([os.path.join(host_build, "test_facade"), "--gpu", prefix], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    W, H, spp, depth = 64, 48, 4, 8
    sc, cam = oracle.cornell_scene(), oracle.Camera(W, H)
    g = np.fromfile(prefix + "_color.bin", np.float32).reshape(-1, 4)
    o, _ = oracle.render(sc, cam, spp, depth, mode=oracle.MODE_FORWARD_FAST)
    ok = ~np.isnan(o[:, :3])
    assert np.array_equal(np.isnan(g[:, :3]), ~ok)
    rel = np.abs(g[:, :3][ok] - o[:, :3][ok]) / np.maximum(np.abs(o[:, :3][ok]), 1e-3 * spp)
    assert (rel <= 1e-4).mean() > 0.9995
    # Camera::CreateRays: bit-exact directions from seeds[i] = i
    dirx = np.fromfile(prefix + "_dirx.bin", np.float32)
    for i in (0, 1, W, W * H - 1, 1234):
        assert dirx[i] == oracle.raygen(cam, i, i)[0][0]
    # MapperPathTracer::intersect on those rays: hit distances and surviving status bits
    t = np.fromfile(prefix + "_t.bin", np.float32)
    status = np.fromfile(prefix + "_status.bin", np.uint8)
    oprim, ot = oracle.primary_hits(sc, cam)
    assert np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    assert np.array_equal(status == 8, oprim >= 0) and set(np.unique(status).tolist()) <= {0, 8}


@pytest.mark.gpu
def test_cli_driver_writes_the_reference_pnm(host_build, tmp_path, oracle):
    """CornellBox_b2pt with the reference's defaults (128x128, 10 spp, depth 5) against the oracle's image,
    compared as 8-bit PNM values (main.cc:339-342: int(255.99*c), unclamped)."""
    out = str(tmp_path / "output")
    r = subprocess.run([os.path.join(host_build, "CornellBox_b2pt"), "-x", "128", "-y", "128", "-samplecount", "10",
                        "-raydepth", "5", "-o", out, "-stats"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "Elapsed time" in r.stdout
    tok = open(out + ".pnm").read().split()
    assert tok[:4] == ["P3", "128", "128", "255"]
    got = np.array(tok[4:], np.int64).reshape(-1, 3)
    o, _ = oracle.render(oracle.cornell_scene(), oracle.Camera(128, 128), 10, 5, mode=oracle.MODE_FORWARD_FAST)
    want = (255.99 * oracle.normalize(o, 10)[:, :3].astype(np.float64)).astype(np.int64)
    assert (np.abs(got - want) <= 1).mean() > 0.9995 and (got == want).mean() > 0.99


@pytest.mark.gpu
def test_cli_hemisphere_sweep_matches_oracle_views(host_build, tmp_path, oracle):
    """-hemisphere (generateHemisphere, main.cc:504-561): one image per (phi, theta) view point, named like the
    reference's generate(); two of the views are checked against the oracle rendered from the same camera."""
    out = str(tmp_path / "view")
    r = subprocess.run([os.path.join(host_build, "CornellBox_b2pt"), "-x", "48", "-y", "48", "-samplecount", "6",
                        "-raydepth", "5", "-hemisphere", "-phicount", "3", "-thetacount", "4", "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "views rendered       = 12" in r.stdout
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("view-") and f.endswith(".pnm"))
    assert len(files) == 12 and "view-0.0000-0.0000.pnm" in files
    rr = np.float32(-1078 / 555.0)
    c = 278 / 555.0
    rphi, rtheta = np.float32(1.0 / 3), np.float32(np.float32(2 * np.pi) / np.float32(4))
    for (iphi, itheta) in [(0, 0), (2, 1)]:
        phi, theta = np.float32(0), np.float32(0)
        for _ in range(iphi):
            phi = np.float32(phi + rphi)
        for _ in range(itheta):
            theta = np.float32(theta + rtheta)
        x = rr * np.cos(theta) * np.sin(phi)
        y = rr * np.sin(theta) * np.sin(phi)
        z = rr * np.cos(phi)
        pos = [np.float32(x + c), np.float32(y + c), np.float32(z + c)]
        name = "view-%.4f-%.4f.pnm" % (phi, theta)
        tok = open(os.path.join(tmp_path, name)).read().split()
        got = np.array(tok[4:], np.int64).reshape(-1, 3)
        o, _ = oracle.render(oracle.cornell_scene(), oracle.Camera(48, 48, pos=pos), 6, 5, mode=oracle.MODE_FORWARD_FAST)
        want = (255.99 * oracle.normalize(o, 6)[:, :3].astype(np.float64)).astype(np.int64)
        assert (np.abs(got - want) <= 1).mean() > 0.999, (name, (np.abs(got - want) <= 1).mean())


@pytest.mark.gpu
def test_cli_fibonacci_sweep_matches_oracle_views(host_build, tmp_path, oracle):
    """-fibonacci (fibonacciHemisphere, main.cc:430-503): the lattice points with z < 0 become camera positions on the
    unit sphere around the box centre, images named output-0.0000-<i>.0000 like generate(cam, ..., 0, i); two views are
    checked against the oracle rendered from the same camera position (float z and radius, double azimuth)."""
    out = str(tmp_path / "fib")
    n, rnd = 12, 5
    r = subprocess.run([os.path.join(host_build, "CornellBox_b2pt"), "-x", "48", "-y", "48", "-samplecount", "6",
                        "-raydepth", "5", "-fibonacci", "-viewcount", str(n), "-viewseed", str(rnd), "-o", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "views rendered       = 6" in r.stdout  # the lower half of the lattice
    files = sorted(f for f in os.listdir(tmp_path) if f.startswith("fib-") and f.endswith(".pnm"))
    assert files == sorted("fib-0.0000-%d.0000.pnm" % i for i in range(6))
    offset = np.float32(2.0 / n)
    inc = np.pi * (3.0 - np.sqrt(5.0))
    c = 278 / 555.0
    for i in (0, 4):
        z = np.float32(np.float32(np.float32(i * offset) - np.float32(1)) + np.float32(offset / np.float32(2)))
        rad = np.float32(np.sqrt(1 - float(z) ** 2))
        phi = ((i + rnd) % n) * inc
        x, y = np.float32(np.cos(phi) * float(rad)), np.float32(np.sin(phi) * float(rad))
        pos = [np.float32(float(x) + c), np.float32(float(y) + c), np.float32(float(z) + c)]
        tok = open(os.path.join(tmp_path, "fib-0.0000-%d.0000.pnm" % i)).read().split()
        got = np.array(tok[4:], np.int64).reshape(-1, 3)
        o, _ = oracle.render(oracle.cornell_scene(), oracle.Camera(48, 48, pos=pos), 6, 5, mode=oracle.MODE_FORWARD_FAST)
        want = (255.99 * oracle.normalize(o, 6)[:, :3].astype(np.float64)).astype(np.int64)
        assert (np.abs(got - want) <= 1).mean() > 0.999, (i, (np.abs(got - want) <= 1).mean())
