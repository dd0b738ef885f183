"""Profiling target for BASELINE configs[3]: 1M random spheres + BVH, 1920x1080, one 16-sample batch, depth 50,
rendered twice.  Used under ncu, never for numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
FLAGS = int(os.environ.get("B2PT_FLAGS", "0"))  # 4096 = B2PT_FLAG_WIDE_BVH
ctx = B.Context(0)
ctx.set_scene(B.Scene.spheres(n)); ctx.build_bvh(FLAGS); ctx.set_camera(B.Camera(1920, 1080))
for rep in range(2):
    ctx.render(16, 50, FLAGS)
    st = ctx.stats()
    print("render %d: %.3f ms, %.3f Gpaths/s, %.3f Gseg/s, launches %d, nodes %d" % (
        rep, st.renderMs, st.paths / st.renderMs / 1e6, st.segments / st.renderMs / 1e6, st.launches, st.bvhNodes))
ctx.close()
