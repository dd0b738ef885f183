"""Measures the non-headline BASELINE.json configs on one B200 (device-timed by the library's CUDA events):
  configs[3]  synthetic 1M random-sphere scene with BVH, 1920x1080, 256 spp, depth 50 (traversal-bound)
  configs[4]  max-depth sweep 1/4/16/50 on the Cornell box 1024^2, 256 spp
Prints one JSON line per measurement.  usage: run_configs.py [spheres|sweep|all] [spp]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B

what = sys.argv[1] if len(sys.argv) > 1 else "all"
FLAGS = int(os.environ.get("B2PT_FLAGS", "0"))  # e.g. 4096 = B2PT_FLAG_WIDE_BVH, 2048 = B2PT_FLAG_ONE_KERNEL_BOUNCE
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 256


def measure(ctx, spp, depth, reps=2):
    best = None
    for _ in range(reps + 1):  # first repetition is the warm-up
        ctx.render(spp, depth, FLAGS)
        st = ctx.stats()
        best = st if best is None or st.renderMs < best.renderMs else best
    return best


if what in ("spheres", "all"):
    n = 1_000_000
    with B.Context(0) as ctx:
        t0 = time.time()
        ctx.set_scene(B.Scene.spheres(n))
        ctx.build_bvh(FLAGS)
        ctx.synchronize()
        build_s = time.time() - t0
        ctx.set_camera(B.Camera(1920, 1080))
        st = measure(ctx, spp, 50)
        print(json.dumps({"config": "1M random spheres + BVH, 1920x1080, %d spp, depth 50" % spp,
                          "path_samples_per_s": st.paths / st.renderMs * 1e3, "segments_per_s": st.segments / st.renderMs * 1e3,
                          "ms": st.renderMs, "segments_per_path": st.segments / st.paths, "bvh_nodes": st.bvhNodes,
                          "scene_upload_plus_bvh_build_s": build_s, "launches": st.launches, "flags": FLAGS}), flush=True)
if what in ("sweep", "all"):
    with B.Context(0) as ctx:
        ctx.set_scene(B.Scene.cornell())
        ctx.build_bvh()
        ctx.set_camera(B.Camera(1024, 1024))
        for depth in (1, 4, 16, 50):
            st = measure(ctx, spp, depth)
            print(json.dumps({"config": "Cornell 1024x1024, %d spp, max depth %d" % (spp, depth),
                              "path_samples_per_s": st.paths / st.renderMs * 1e3,
                              "segments_per_s": st.segments / st.renderMs * 1e3, "ms": st.renderMs,
                              "segments_per_path": st.segments / st.paths, "launches": st.launches}), flush=True)
