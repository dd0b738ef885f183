"""Time the hemisphere sweep (the reference's dataset-generation use, main.cc:431-561) two ways on cuda:0:
  loop    one b2pt_set_camera + b2pt_render + b2pt_read_color per view (what a port of the reference's loop does)
  views   one b2pt_render_views call for the whole list (view-batched launches, one D2H)
  pnm16   the same with B2PT_FLAG_VIEWS_PNM16 (the P3 writer's integers packed on the GPU, 6 B/pixel to the host)
Prints one JSON line per canvas configuration.  Host wall clock around complete calls (results on the host)."""
import json
import math
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B  # noqa: E402


def views_on_hemisphere(n_phi, n_theta):
    c = np.array([278 / 555.0] * 3, np.float32)
    r = 1078 / 555.0
    out = []
    for a in range(n_phi):
        phi = (a + 0.5) / n_phi
        for b in range(n_theta):
            th = 2 * math.pi * b / n_theta
            pos = c + r * np.array([math.sin(phi) * math.cos(th), math.cos(phi), math.sin(phi) * math.sin(th)])
            out.append(np.concatenate([pos, c, [0, 1, 0], [40.0]]).astype(np.float32))
    return np.stack(out)


def main():
    ctx = B.Context(0)
    ctx.set_scene(B.Scene.cornell())
    ctx.build_bvh()
    for (W, H, spp, depth, nphi, nth) in [(128, 128, 10, 5, 15, 15), (128, 128, 10, 50, 15, 15),
                                          (256, 256, 64, 50, 8, 8), (512, 512, 256, 50, 2, 4)]:
        views = views_on_hemisphere(nphi, nth)
        V = views.shape[0]
        out = np.empty((V, W * H, 4), np.float32)
        res = {}
        for rep in range(3):  # first repetition warms allocations up
            t0 = time.perf_counter()
            for k, v in enumerate(views):
                ctx.set_camera(B.Camera(W, H, pos=v[0:3], lookAt=v[3:6], up=tuple(v[6:9]), fov=float(v[9])))
                ctx.render(spp, depth)
                ctx.read_color(out[k])
            res["loop"] = time.perf_counter() - t0
        loop_img = out.copy()
        for rep in range(3):
            t0 = time.perf_counter()
            got = ctx.render_views(views, W, H, spp, depth, out=out)
            res["views"] = time.perf_counter() - t0
        st = ctx.stats()
        out16 = np.empty((V, W * H, 3), np.uint16)
        for rep in range(3):
            t0 = time.perf_counter()
            ctx.render_views(views, W, H, spp, depth, flags=B.FLAG_VIEWS_PNM16, out=out16)
            res["pnm16"] = time.perf_counter() - t0
        same = bool(np.array_equal(got.view(np.uint32), loop_img.view(np.uint32)))
        print(json.dumps({"canvas": [W, H], "spp": spp, "depth": depth, "views": V,
                          "loop_views_per_s": V / res["loop"], "batched_views_per_s": V / res["views"],
                          "batched_pnm16_views_per_s": V / res["pnm16"],
                          "loop_paths_per_s": V * W * H * spp / res["loop"],
                          "batched_paths_per_s": V * W * H * spp / res["views"],
                          "speedup": res["loop"] / res["views"], "bit_identical": same,
                          "batches": st.batches, "launches": st.launches}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
