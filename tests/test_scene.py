"""Host-side scene builders of the product (csrc/b2pt_scene.cpp) against the oracle's independent restatement."""
import numpy as np


FIELDS = ("pts", "quadIds", "sphPt", "sphR", "matIdxQ", "texIdxQ", "matIdxS", "texIdxS", "matType", "texType", "tex",
          "lightQuadIds", "lightSphPt", "lightSphR")


def test_cornell_builder_is_bit_identical_to_the_oracle(b2pt, oracle):
    a, b = b2pt.Scene.cornell(), oracle.cornell_scene()
    for k in FIELDS:
        x, y = getattr(a, k), getattr(b, k)
        assert x.dtype == y.dtype and x.shape == y.shape, k
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8)), k
    assert a.lightables == b.lightables == 2 and a.refIdx == b.refIdx == 1.5


def test_spheres_scene_definition(b2pt, oracle):
    """BASELINE.json configs[3]: sphere k from a private wang stream seeded with k."""
    n = 1000
    s = b2pt.Scene.spheres(n)
    assert s.pts.shape == (n + 8, 3) and s.quadIds.shape == (2, 5) and len(s.sphR) == n
    for k in (0, 1, 17, 999):
        ra, rb, rc, rd = (np.float32(x) for x in oracle.randf_chain(k, 4))
        assert np.array_equal(s.pts[k], np.array([ra, np.float32(0.5) * rb, rc], np.float32))
        assert s.sphR[k] == np.float32(0.002) * (np.float32(0.5) + rd)
        assert s.matIdxS[k] == k % 3 and s.texIdxS[k] == k % 3
    assert np.allclose(s.pts[n:n + 4, 1], 0.98) and np.allclose(s.pts[n + 4:, 1], 0.0)
    assert s.quadIds.tolist() == [[0, n, n + 1, n + 2, n + 3], [1, n + 4, n + 5, n + 6, n + 7]]
    assert s.matIdxQ.tolist() == [3, 1] and s.lightQuadIds.tolist() == [[0, n, n + 1, n + 2, n + 3]]
    assert s.lightSphPt.tolist() == [0] and s.lightSphR[0] == s.sphR[0]


def test_scene_builder_rejects_null(b2pt):
    assert b2pt.lib().b2pt_scene_cornell(*([None] * 11)) == b2pt.ERR_BAD_VALUE
    assert b2pt.lib().b2pt_scene_spheres(0, *([None] * 11)) == b2pt.ERR_BAD_VALUE
