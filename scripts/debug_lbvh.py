import sys, time, numpy as np
sys.path.insert(0,'/root/repo')
import raytracingtherestofyourlife_b200 as B
n,W,H=1000000,960,540
s=B.Scene.spheres(n)
with B.Context(0) as ctx:
    ctx.set_scene(s); ctx.set_camera(B.Camera(W,H))
    t=time.time(); ctx.build_bvh(0); ctx.synchronize(); print("host build", time.time()-t)
    p0,t0=ctx.primary_hits()
    ctx.render(2,50,0); a=ctx.read_color(); sa=ctx.stats()
    t=time.time(); ctx.build_bvh(B.FLAG_GPU_LBVH); ctx.synchronize(); print("gpu build", time.time()-t)
    t=time.time(); ctx.build_bvh(B.FLAG_GPU_LBVH); ctx.synchronize(); print("gpu build again", time.time()-t)
    p1,t1=ctx.primary_hits()
    ctx.render(2,50,B.FLAG_GPU_LBVH); b=ctx.read_color(); sb=ctx.stats()
    d=(p0!=p1); print("prim diffs", d.sum(), "t diffs", (t0.view(np.uint32)!=t1.view(np.uint32)).sum())
    idx=np.nonzero(d)[0][:10]
    for i in idx: print(i, p0[i], p1[i], t0[i], t1[i])
    print("segments", sa.segments, sb.segments, "img equal", np.array_equal(a,b,equal_nan=True), "ndiff px", (a!=b).any(1).sum(), sa.renderMs, sb.renderMs)
