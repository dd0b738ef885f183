"""Histogram of the small-scene trace's two-phase filter (experiment build: scripts/build_variant.sh hist
-DB2PT_DEBUG_HIST, run with B2PT_LIB=variants/libb2pt_hist.so): candidates per ray after phase 1 and exact
Lagae-Dutre tests per ray in phase 2, for primary rays (depth 1) and for all bounces (depth 50)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import raytracingtherestofyourlife_b200 as B
L = B.lib()
ctx = B.Context(0)
ctx.set_scene(B.Scene.cornell()); ctx.build_bvh(); ctx.set_camera(B.Camera(512, 512))
buf = (C.c_ulonglong * 256)()
for name, depth, flags in (("primary rays (generic filter)", 1, B.FLAG_NO_PRIMARY_MASKS), ("all bounces, depth 50", 50, 0)):
    L.b2pt_debug_hist(None, 1)
    ctx.render(16, depth, flags | B.FLAG_NO_OVERLAP)
    ctx.synchronize()
    L.b2pt_debug_hist(buf, 0)
    h = list(buf)
    n = sum(h[0:16])
    print(json.dumps({"what": name, "rays": n, "candidates_per_ray": [round(x / max(n, 1), 4) for x in h[0:16]],
                      "exact_tests_per_ray": [round(x / max(n, 1), 4) for x in h[16:32]],
                      "mean_candidates": sum(i * x for i, x in enumerate(h[0:16])) / max(n, 1),
                      "mean_exact_tests": sum(i * x for i, x in enumerate(h[16:32])) / max(n, 1),
                      "filtered_quads_hit_fraction": h[33] / max(h[32] + h[33], 1)}))
L.b2pt_debug_hist(None, 1)
ctx.set_camera(B.Camera(1024, 1024))
ctx.render(4, 1, B.FLAG_NO_OVERLAP)
ctx.synchronize()
L.b2pt_debug_hist(buf, 0)
h = list(buf)
print(json.dumps({"what": "primary tiles at 1024x1024 (masked path)", "tiles": h[40], "filter_candidates_per_tile": h[41] / max(h[40], 1),
                  "gate_bits_per_tile": h[42] / max(h[40], 1), "tiles_with_gate_bits": h[43] / max(h[40], 1),
                  "tiles_with_0_1_2_3plus_candidates": [round(x / max(h[40], 1), 4) for x in h[44:48]]}))
L.b2pt_debug_hist(None, 1)
ctx.set_camera(B.Camera(512, 512))
ctx.render(16, 50, B.FLAG_NO_OVERLAP | B.FLAG_NO_PRIMARY_MASKS)
ctx.synchronize()
L.b2pt_debug_hist(buf, 0)
h = list(buf)
n = max(sum(h[0:16]), 1)
print(json.dumps({"what": "phase 2 by visit index (all bounces): share of rays", "rays": n,
                  "first_test": {v: round(h[64 + v] / n, 4) for v in range(32) if h[64 + v]},
                  "first_test_failed": {v: round(h[96 + v] / n, 4) for v in range(32) if h[96 + v]},
                  "later_tests": {v: round(h[128 + v] / n, 4) for v in range(32) if h[128 + v]},
                  "later_tests_won": {v: round(h[160 + v] / n, 4) for v in range(32) if h[160 + v]},
                  "first_hit_but_loop_continued": {v: round(h[192 + v] / n, 4) for v in range(32) if h[192 + v]}}))
ctx.close()
