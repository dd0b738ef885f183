"""Profiling target for the tail mode: Cornell 1024x1024, 64 spp = two 32-sample batches, depth 50 (the second batch
switches to the global-queue tail kernels).  Used under ncu, never for numbers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
ctx = B.Context(0)
ctx.set_scene(B.Scene.cornell()); ctx.build_bvh(); ctx.set_camera(B.Camera(1024, 1024))
ctx.render(64, 50, int(sys.argv[1]) if len(sys.argv) > 1 else 0)
st = ctx.stats()
print("render: %.3f ms, tailDepth %d, launches %d" % (st.renderMs, st.tailDepth, st.launches))
ctx.close()
