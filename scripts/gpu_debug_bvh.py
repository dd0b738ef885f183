import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracingtherestofyourlife_b200 as B
from oracle import oracle as O
ctx = B.Context(0); ctx.set_scene(B.Scene.cornell()); ctx.build_bvh()
cam = B.Camera(128, 96); ctx.set_camera(cam)
ctx.render(1, 1, 0); p0, t0 = ctx.primary_hits()
ctx.render(1, 1, B.FLAG_FORCE_BVH); st = ctx.stats(); p1, t1 = ctx.primary_hits()
op, ot = O.primary_hits(O.cornell_scene(), O.Camera(128, 96))
print("bvh nodes", st.bvhNodes, "traced quads", st.tracedQuads)
print("small vs oracle mism", (p0 != op).sum(), "bvh vs oracle mism", (p1 != op).sum())
d = np.argwhere(p1 != op).ravel()
print(d[:20]); print("oracle", op[d][:20]); print("bvh   ", p1[d][:20]); print("t orc", ot[d][:8], "t bvh", t1[d][:8])
u, c = np.unique(op[d], return_counts=True); print("oracle prims at mismatches", dict(zip(u.tolist(), c.tolist())))
u, c = np.unique(p1[d], return_counts=True); print("bvh prims at mismatches", dict(zip(u.tolist(), c.tolist())))
