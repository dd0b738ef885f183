"""Small target that runs every kernel family of the library once (written for `compute-sanitizer --tool memcheck`, which
this pool has closed; it still serves as a quick all-paths run), at sizes that finish
in seconds under the sanitizer (Cornell 64x48 and a 300-sphere BVH scene).  Not a test of results (tests/ do that)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracingtherestofyourlife_b200 as B

os.environ["B2PT_BATCH_PATHS"] = str(64 * 48 * 2)  # several batches in flight: tail modes and the cluster loop take part
os.environ["B2PT_TAIL_RAYS_PER_WARP"] = "100000"
cam = B.Camera(64, 48)
with B.Context(0) as ctx:
    ctx.set_scene(B.Scene.cornell()); ctx.build_bvh(); ctx.set_camera(cam)
    for flags in (0, B.FLAG_ONE_KERNEL_BOUNCE, B.FLAG_NO_PRIMARY_MASKS, B.FLAG_NO_TAIL, B.FLAG_REFERENCE_STREAM,
                  B.FLAG_ONE_KERNEL_BOUNCE | B.FLAG_REFERENCE_STREAM, B.FLAG_FORCE_BVH, B.FLAG_FORCE_BVH | B.FLAG_WIDE_BVH):
        ctx.render(8, 12, flags)
        print("cornell flags", flags, "segments", ctx.stats().segments, flush=True)
    ctx.render(2, 6, 0); ctx.primary_hits(); ctx.render_direct(); ctx.read_pnm16(8)
    views = np.array([[0.5 + 2 * np.cos(t), 0.5, 0.5 + 2 * np.sin(t), 0.5, 0.5, 0.5, 0, 1, 0, 40.0] for t in (0.5, 2.0, 4.0)], np.float32)
    ctx.render_views(views, 32, 32, 4, 5)
    ctx.render_views(views, 32, 32, 4, 5, flags=B.FLAG_VIEWS_PNM16)
for flags in (0, B.FLAG_WIDE_BVH, B.FLAG_GPU_LBVH, B.FLAG_NO_RAY_SORT, B.FLAG_SPLIT_TRACE, B.FLAG_GPU_LBVH | B.FLAG_WIDE_BVH):
    with B.Context(0) as ctx:
        s = B.Scene.spheres(300)
        ctx.set_scene(s); ctx.build_bvh(flags); ctx.set_camera(cam)
        ctx.render(6, 10, flags)
        print("spheres flags", flags, "segments", ctx.stats().segments, flush=True)
        if not (flags & B.FLAG_WIDE_BVH):
            ctx.update_spheres(s.pts[:300] + np.float32(0.01)); ctx.refit_bvh(); ctx.render(2, 6, flags)
        ctx.primary_hits()
print("sanitize target done")
