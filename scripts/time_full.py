"""A/B timing target: BASELINE.json configs[1] (Cornell 1024x1024, 1024 spp, depth 50), three renders after a warm-up.
Library selected with B2PT_LIB (scripts/time_variants.sh).  Prints the best and the median GPU render time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
import numpy as np
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # e.g. 1024 = B2PT_FLAG_NO_PRIMARY_MASKS
ctx = B.Context(0)
ctx.set_scene(B.Scene.cornell()); ctx.build_bvh(); ctx.set_camera(B.Camera(1024, 1024))
ms = []
for rep in range(5):
    ctx.render(spp, 50, flags)
    st = ctx.stats()
    ms.append(st.renderMs)
img = ctx.read_color()
print("renders ms %s  best %.2f median %.2f  Gpaths/s(best) %.3f  segments %d  checksum %.6e" % (
    ["%.1f" % m for m in ms], min(ms[1:]), float(np.median(ms[1:])), st.paths / min(ms[1:]) / 1e6, st.segments,
    float(np.nansum(img[:, :3].astype(np.float64)))))
ctx.render(128, 50, flags | B.FLAG_NO_OVERLAP)  # one batch, launches serialised: per-launch CUDA events
print("stage profile (trace ms, shade ms, rays in), bounces 0..5:", [tuple(round(x, 3) if isinstance(x, float) else x for x in e)
                                                                      for e in ctx.stage_profile(6)])
ctx.close()
