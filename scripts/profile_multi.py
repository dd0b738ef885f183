"""Timing target with several batches in flight: Cornell 1024x1024, SPP spp (default 256), depth 50, three renders."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flags = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ctx = B.Context(0)
ctx.set_scene(B.Scene.cornell()); ctx.build_bvh(); ctx.set_camera(B.Camera(1024, 1024))
best = 1e9
for rep in range(3):
    ctx.render(spp, 50, flags)
    best = min(best, ctx.stats().renderMs)
st = ctx.stats()
print("best of 3: %.3f ms, %.3f Gpaths/s, %.3f Gseg/s, batches %d" % (best, st.paths / best / 1e6, st.segments / best / 1e6, st.batches))
ctx.close()
