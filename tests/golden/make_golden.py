"""Generates the committed golden fixtures of tests/golden/.

Provenance of each group:
  "survey"      derived independently by the surveyor's throwaway numpy probe.
  "refworklets" outputs of the REFERENCE'S OWN header-only worklets and scene builder (CornellBox.cpp), compiled from /root/reference against the
                VTK-m stand-in of oracle/vtkm_min/ and driven in the reference's launch order by
                oracle/ref_harness.cxx (only generated when /root/reference is present; the reference as a whole
                needs VTK-m and cannot be built here).  These pin the C oracle to the reference's code.
  "oracle"      outputs of the C oracle (oracle/b2pt_oracle.c) itself.
Run:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def main():
    sc = O.cornell_scene()
    g = {
        "survey": {  # SURVEY.md 8c, independent derivation
            "wang_chain": {"0": [3232319850, 3075307816, 755367838, 413455686],
                           "1": [663891101, 1738326990, 801461103, 3205955024],
                           "2": [3329832309, 685338552, 3175962347, 68000511]},
            "randf_seed0": [0.75258309, 0.71602589, 0.17587277, 0.09626515],
            "wang_init_0_3": [413455686, 3205955024, 68000511, 2223342941],
            "primary_hist_64": {"-1": 379, "0": 606, "1": 568, "2": 29, "3": 572, "4": 166, "5": 813, "7": 294,
                                "8": 28, "10": 3, "12": 139, "13": 5, "14": 19, "17": 430, "19": 45},
        },
        "oracle": {},
    }
    o = g["oracle"]
    o["wang_chain_12345"] = O.wang_chain(12345, 8)
    o["scene_pts_sha256"] = hashlib.sha256(sc.pts.tobytes()).hexdigest()
    o["scene_quadIds_sha256"] = hashlib.sha256(sc.quadIds.tobytes()).hexdigest()
    for W in (64, 128):
        prim, t = O.primary_hits(sc, O.Camera(W, W))
        np.save(os.path.join(HERE, "primary_ids_%d.npy" % W), prim.astype(np.int8))
        o["primary_t_sha256_%d" % W] = hashlib.sha256(t.tobytes()).hexdigest()
    # BASELINE.json configs[0]: 128x128, 10 spp, depth 5, reference-faithful pass-per-worklet mode
    img, st = O.render(sc, O.Camera(128, 128), 10, 5, mode=O.MODE_PASSES)
    np.save(os.path.join(HERE, "config1_passes_rgb.npy"), img[:, :3].copy())
    o["config1"] = {"segments": int(st.segments), "rngDraws": int(st.rngDraws), "nanSamples": int(st.nanSamples),
                    "alive": [int(st.aliveAtDepth[k]) for k in range(5)]}
    img3, st3 = O.render(sc, O.Camera(128, 128), 10, 5, mode=O.MODE_FORWARD_FAST)
    np.save(os.path.join(HERE, "config1_fast_rgb.npy"), img3[:, :3].copy())
    o["config1_fast"] = {"segments": int(st3.segments), "nanSamples": int(st3.nanSamples)}
    cam = O.Camera(128, 128)
    a, b, c = cam.basis()
    o["camera_basis_128"] = {"nlook": a.view(np.uint32).tolist(), "dx": b.view(np.uint32).tolist(),
                             "dy": c.view(np.uint32).tolist()}
    from oracle import refharness as R
    if R.available():
        r = g["refworklets"] = {}
        r["wang_chain_12345"] = R.wang_chain(12345, 8)
        r["randf_seed0_bits"] = np.array(R.randf_chain(0, 8), np.float32).view(np.uint32).tolist()
        # BASELINE.json configs[0] through the reference's worklets (median-split stand-in BVH)
        rimg, rseg, rt0, rhit0 = R.render(sc, O.Camera(128, 128), 10, 5)
        np.save(os.path.join(HERE, "config1_refworklets_rgb.npy"), rimg[:, :3].copy())
        r["config1"] = {"segments": rseg, "t0_sha256": hashlib.sha256(rt0.tobytes()).hexdigest(),
                        "hit0_sha256": hashlib.sha256(rhit0.tobytes()).hexdigest()}
        # deep paths: 64x64, 8 spp, depth 50
        rimg, rseg, rt0, rhit0 = R.render(sc, O.Camera(64, 64), 8, 50)
        np.save(os.path.join(HERE, "deep64_refworklets_rgb.npy"), rimg[:, :3].copy())
        r["deep64"] = {"segments": rseg, "t0_sha256": hashlib.sha256(rt0.tobytes()).hexdigest()}
        # primary-ray closest distances at 256x256 (45 rays depend on the leaf-box gate, 1 on tree shape)
        _, _, rt0, rhit0 = R.render(sc, O.Camera(256, 256), 1, 1)
        np.save(os.path.join(HERE, "primary_t_256_refworklets.npy"), rt0)
        # the reference's own scene builder (CornellBox.cpp compiled where it lies, oracle/ref_scene.cxx): bit patterns
        rs = R.cornell_scene()
        r["cornell_scene"] = {k: (v.view(np.uint32) if v.dtype == np.float32 else v).ravel().tolist()
                              for k, v in rs.items()}
    else:
        with open(os.path.join(HERE, "golden.json")) as f:
            old = json.load(f)
        if "refworklets" in old:
            g["refworklets"] = old["refworklets"]
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()
