"""Traversal statistics of the BVH kernels on the 1M-sphere scene (experiment build: scripts/build_variant.sh hist
-DB2PT_DEBUG_HIST, run with B2PT_LIB=variants/libb2pt_hist.so): node visits, children hit and exact primitive tests per
ray for the 8-wide tree and for the binary tree (all bounces of a 4-spp render)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
L = B.lib()
buf = (C.c_ulonglong * 256)()
for name, flags in (("8-wide tree", B.FLAG_WIDE_BVH), ("binary tree", 0)):
    ctx = B.Context(0)
    ctx.set_scene(B.Scene.spheres(1000000)); ctx.build_bvh(flags); ctx.set_camera(B.Camera(1920, 1080))
    L.b2pt_debug_hist(None, 1)
    ctx.render(4, 50, flags | B.FLAG_NO_OVERLAP)
    ctx.synchronize()
    L.b2pt_debug_hist(buf, 0)
    h = list(buf)
    st = ctx.stats()
    if flags:
        r = max(h[42], 1)
        print(json.dumps({"tree": name, "rays": h[42], "segments": st.segments, "node_visits_per_ray": h[40] / r,
                          "inner_children_hit_per_visit": h[43] / max(h[40], 1), "prim_boxes_hit_per_visit": h[44] / max(h[40], 1),
                          "exact_prim_tests_per_ray": h[41] / r}))
    else:
        r = max(h[51], 1)
        print(json.dumps({"tree": name, "rays": h[51], "segments": st.segments, "child_pair_steps_per_ray": h[48] / r,
                          "leaves_per_ray": h[49] / r, "prims_offered_per_ray": h[50] / r}))
    ctx.close()
