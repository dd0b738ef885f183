"""Generates the committed golden fixtures of tests/golden/ from the CPU oracle.

Provenance: the reference ships no tests or golden vectors and cannot be built here (VTK-m absent), so these
are ORACLE outputs ("parity unpinned").  The entries under "survey" were derived independently by the
surveyor's throwaway numpy probe and are the only vectors not produced by this oracle.
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
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(g, f, indent=1)
    print("wrote golden fixtures to", HERE)


if __name__ == "__main__":
    main()
