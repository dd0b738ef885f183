"""One reproducible record per BASELINE.json configuration that bench.py's headline line does not cover, with the same
schema (value, e2e, roofline, cpu_baseline, clocks), on ONE B200:

  configs[0]  CornellBox.cpp scene at the reference's defaults (128x128, 10 spp, depth 5; main.cc:56-62) -- the
              reference's own CPU-runnable case: the CPU arm runs the whole config, all cores and one thread
  configs[3]  synthetic 1M random-sphere scene with BVH, 1920x1080, 256 spp, depth 50
  configs[4]  max-depth sweep 1 / 4 / 16 / 50 on the Cornell box 1024^2, 256 spp

(configs[1] is bench.py itself; configs[2], 4096^2 x 4096 spp sharded over 2/4/8 GPUs, is
`torchrun ... bench.py --gpus N --size 4096 --total-spp 4096`.)  Prints one JSON line per record.

usage: python scripts/bench_configs.py [c0] [c3] [c4]     (default: all)"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
import raytracingtherestofyourlife_b200 as B  # noqa: E402

which = set(sys.argv[1:]) or {"c0", "c3", "c4"}
PEAK, PEAK_SRC = bench.measured_peak()
CORES = os.cpu_count() or 1


def gpu_record(label, scene, cam, spp, depth, steps, warmup, build_flags=0):
    """Device-timed renders with everything resident (value) and host-buffer renders through the C-ABI (e2e)."""
    N = cam.W * cam.H
    ctx = B.Context(0)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    t0 = time.perf_counter()
    ctx.set_scene(scene)
    ctx.build_bvh(build_flags)
    ctx.synchronize()
    build_s = time.perf_counter() - t0
    ctx.set_camera(cam)
    for _ in range(warmup):
        ctx.render(spp, depth, build_flags)
    torch.cuda.synchronize()
    sampler = bench.ClockSampler(0)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(steps):
        ctx.render(spp, depth, build_flags)
    ev1.record(stream)
    torch.cuda.synchronize()
    sampler.stop_flag.set()
    sampler.join()
    ms = ev0.elapsed_time(ev1) / steps
    st = ctx.stats()
    host = torch.empty((N, 4), dtype=torch.float32).pin_memory()

    def e2e_step():
        ctx.set_scene(scene)
        ctx.build_bvh(build_flags)
        ctx.set_camera(cam)
        ctx.render(spp, depth, build_flags)
        ctx.read_color(host.data_ptr())

    e2e_step()
    t0 = time.perf_counter()
    e_steps = max(1, min(steps, 3))
    for _ in range(e_steps):
        e2e_step()
    e_dt = (time.perf_counter() - t0) / e_steps
    paths = float(N) * spp
    gbs = st.segments * bench.ALGO_BYTES_PER_SEGMENT / (ms * 1e-3) / 1e9
    line = {
        "metric": "path_samples_per_s", "value": paths / (ms * 1e-3), "unit": bench.UNIT, "n_gpus": 1, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": label, "segments_per_path": st.segments / paths, "batches_per_step": st.batches,
                   "trace_path": "BVH (8-wide compressed / binary)" if st.tracePath else "small scene (kernel parameter)",
                   "scene_upload_plus_structure_build_s": build_s,
                   "l2": "inputs are generated on the device; the records in flight (272 B per path) exceed the 126 MB "
                         "L2 whenever a batch holds more than 0.5 M paths"},
        "segments_per_s": st.segments / (ms * 1e-3),
        "roofline": {"bound": "hbm", "achieved": gbs, "peak": PEAK, "unit": "GB/s", "frac": gbs / PEAK, "traffic": None,
                     "peak_source": PEAK_SRC, "kernel": "all bounce launches of a step (88 B per live segment)",
                     "algorithmic_bytes_per_segment": bench.ALGO_BYTES_PER_SEGMENT},
        "e2e": {"value": paths / e_dt, "unit": bench.UNIT, "h2d_bytes_per_step": int(scene.nbytes()),
                "d2h_bytes_per_step": int(N * 16)},
        "gpu_launches": int(st.launches) * steps, "clocks": sampler.summary(),
    }
    ctx.close()
    return line


def cpu_reference(W, H, spp, depth, threads):
    """The reference's own worklets (oracle/_ref timing build) on `threads` host threads: paths/s, seconds."""
    kind, desc, step, build = bench.cpu_reference_arm()
    assert W == H
    val, dt, _ = step(W, spp, depth, threads)
    return kind, val, dt, build[0], desc


if "c0" in which:
    line = gpu_record("BASELINE.json configs[0]: CornellBox.cpp scene, 128x128, 10 spp, max depth 5 (main.cc:56-62)",
                      B.Scene.cornell(), B.Camera(128, 128), 10, 5, steps=50, warmup=5)
    kind, v_all, dt_all, build, desc = cpu_reference(128, 128, 10, 5, CORES)
    _, v_one, dt_one, _, _ = cpu_reference(128, 128, 10, 5, 1)
    line["cpu_baseline"] = {"value": v_all, "unit": bench.UNIT, "cores": CORES, "kind": kind, "one_thread_value": v_one,
                            "sample": "the whole config: %.3f s on %d threads, %.3f s on one thread; built %s; %s" % (
                                dt_all, CORES, dt_one, build, desc)}
    print(json.dumps(line), flush=True)

if "c4" in which:
    for depth in (1, 4, 16, 50):
        line = gpu_record("BASELINE.json configs[4]: Cornell box 1024x1024, 256 spp, max depth %d" % depth,
                          B.Scene.cornell(), B.Camera(1024, 1024), 256, depth, steps=5, warmup=3)
        kind, v_all, dt_all, build, desc = cpu_reference(1024, 1024, 1, depth, CORES)
        line["cpu_baseline"] = {"value": v_all, "unit": bench.UNIT, "cores": CORES, "kind": kind,
                                "sample": "1024x1024, depth %d, 1 of 256 spp in %.1f s (cost is linear in spp); built %s"
                                          % (depth, dt_all, build)}
        print(json.dumps(line), flush=True)

if "c3" in which:
    n = 1_000_000
    scene = B.Scene.spheres(n)
    line = gpu_record("BASELINE.json configs[3]: synthetic 1M random-sphere scene with BVH, 1920x1080, 256 spp, max depth 50",
                      scene, B.Camera(1920, 1080), 256, 50, steps=2, warmup=1)
    # no reference counterpart (SphereExtractor.cxx:108-111 is only valid for one sphere): the C restatement with its
    # own brute-force-checked traversal, on a bounded sample
    from oracle import oracle as O
    O.build()
    osc = O.Scene(scene.pts, scene.quadIds, scene.sphPt, scene.sphR, scene.matIdxQ, scene.texIdxQ, scene.matIdxS,
                  scene.texIdxS, scene.matType, scene.texType, scene.tex, scene.lightQuadIds, scene.lightSphPt,
                  scene.lightSphR, scene.lightables, scene.refIdx)
    t0 = time.perf_counter()
    _, ost = O.render(osc, O.Camera(64, 36), 1, 50, mode=O.MODE_FORWARD_FAST)  # brute force over 1M spheres per segment
    dt = time.perf_counter() - t0
    line["cpu_baseline"] = {"value": ost.paths / dt, "unit": bench.UNIT, "cores": CORES, "kind": "port",
                            "sample": "64x36, 1 spp, depth 50 (%d paths, %d segments) in %.1f s: the oracle's early-exit "
                                      "mode, brute force over all primitives (it is the checker, it has no tree); the "
                                      "reference has no multi-sphere path" % (
                                          ost.paths, ost.segments, dt)}
    print(json.dumps(line), flush=True)
