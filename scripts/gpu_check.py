"""Ad-hoc first GPU check: parity vs the oracle and a quick timing. Writes gpurun_out/gpu_check.log."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import raytracingtherestofyourlife_b200 as B
from oracle import oracle as O

def cmp(name, g, o, spp):
    g3, o3 = g[:, :3].astype(np.float64), o[:, :3].astype(np.float64)
    den = np.maximum(np.abs(o3), 1e-3 * spp)
    rel = np.abs(g3 - o3) / den
    pix = rel.max(1)
    print("%s: mean gpu %s oracle %s | max rel %.3e | pixels within 1e-4: %.5f  within 1e-2: %.5f | exact-equal px %.5f" % (
        name, (g3.mean(0) / spp).round(5), (o3.mean(0) / spp).round(5), pix.max(), (pix < 1e-4).mean(), (pix < 1e-2).mean(),
        (g3 == o3).all(1).mean()), flush=True)

ctx = B.Context(0)
sc, osc = B.Scene.cornell(), O.cornell_scene()
ctx.set_scene(sc); ctx.build_bvh()
for W in (128, 1024):
    cam = B.Camera(W, W); ctx.set_camera(cam)
    t0 = time.time(); gp, gt = ctx.primary_hits(); t1 = time.time()
    op, ot = O.primary_hits(osc, O.Camera(W, W))
    print("primary %dx%d: id mismatches %d / %d ; t bit-equal %s ; gpu %.3fs" % (W, W, int((gp != op).sum()), gp.size,
          np.array_equal(gt.view(np.uint32), ot.view(np.uint32)), t1 - t0), flush=True)

cam = B.Camera(128, 128); ocam = O.Camera(128, 128); ctx.set_camera(cam)
ctx.render(10, 5, B.FLAG_REFERENCE_STREAM); g = ctx.read_color(); st = ctx.stats()
o2, os2 = O.render(osc, ocam, 10, 5, mode=O.MODE_FORWARD_BURN)
o0, os0 = O.render(osc, ocam, 10, 5, mode=O.MODE_PASSES)
cmp("cfg1 refstream vs oracle FORWARD_BURN", g, o2, 10)
cmp("cfg1 refstream vs oracle PASSES", g, o0, 10)
print("  segments gpu %d oracle %d nan %d/%d launches %d ms %.3f" % (st.segments, os2.segments, st.nanSamples, os2.nanSamples, st.launches, st.renderMs))
ctx.render(10, 5, 0); g = ctx.read_color(); st = ctx.stats()
o3, os3 = O.render(osc, ocam, 10, 5, mode=O.MODE_FORWARD_FAST)
cmp("cfg1 fast vs oracle FORWARD_FAST", g, o3, 10)
print("  segments gpu %d oracle %d" % (st.segments, os3.segments))
ctx.render(64, 50, 0); g = ctx.read_color(); st = ctx.stats()
o3, os3 = O.render(osc, ocam, 64, 50, mode=O.MODE_FORWARD_FAST)
cmp("128^2 64spp D50 fast vs oracle FORWARD_FAST", g, o3, 64)
print("  segments gpu %d oracle %d" % (st.segments, os3.segments))
ctx.render(64, 50, B.FLAG_FORCE_BVH); g2 = ctx.read_color(); st = ctx.stats()
cmp("128^2 64spp D50 fast BVH-path vs oracle", g2, o3, 64)
print("  bvh nodes %d tracePath %d equal-to-small %s" % (st.bvhNodes, st.tracePath, np.array_equal(g, g2)))

for W, spp in ((1024, 16), (1024, 64), (1024, 256)):
    cam = B.Camera(W, W); ctx.set_camera(cam)
    for rep in range(2):
        ctx.render(spp, 50, 0); st = ctx.stats()
    print("timing %dx%d spp %d D50: %.2f ms, %.3f Gpaths/s, %.3f Gseg/s, seg/path %.3f, batches %d x %d, launches %d" % (
        W, W, spp, st.renderMs, st.paths / st.renderMs / 1e6, st.segments / st.renderMs / 1e6, st.segments / st.paths, st.batches, st.samplesPerBatch, st.launches), flush=True)
ctx.render(64, 50, B.FLAG_KILL_ZERO_THROUGHPUT); st = ctx.stats()
print("kill-zero: %.2f ms %.3f Gpaths/s seg/path %.3f" % (st.renderMs, st.paths / st.renderMs / 1e6, st.segments / st.paths))
