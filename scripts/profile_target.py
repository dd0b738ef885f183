"""Short profiling target: Cornell 1024x1024, one 128-sample batch (the batch the bench's steps are made of), depth 50, rendered twice (warm-up + measured).
Each render = 50 k_bounce launches + 1 k_accumulate.  Used under ncu (launch list / --set full), never for numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
flags = int(sys.argv[1]) if len(sys.argv) > 1 else B.FLAG_NO_OVERLAP  # one batch, launches serialised
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 128
ctx = B.Context(0)
ctx.set_scene(B.Scene.cornell()); ctx.build_bvh(); ctx.set_camera(B.Camera(1024, 1024))
for rep in range(2):
    ctx.render(spp, 50, flags)
    st = ctx.stats()
    print("render %d: %.3f ms, %.3f Gpaths/s, %.3f Gseg/s, launches %d" % (
        rep, st.renderMs, st.paths / st.renderMs / 1e6, st.segments / st.renderMs / 1e6, st.launches))
ctx.close()
