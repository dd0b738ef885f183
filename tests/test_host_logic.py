"""Host-side logic of the library that needs no GPU: the batch planner behind b2pt_render / b2pt_render_views."""
import pytest


def test_plan_batches_rules(b2pt):
    Mi = 1 << 20
    plan = b2pt.plan_batches
    N = 1024 * 1024
    # the measured optima on the 1024^2 canvas: four batches in flight unless they drop below 32 Mi paths
    assert plan(64, N) == (32, 2)
    assert plan(128, N) == (32, 4)
    assert plan(256, N) == (64, 4)
    assert plan(512, N) == (128, 4)
    assert plan(1024, N) == (128, 8)       # memory target: 128 Mi paths per batch
    assert plan(1000, N) == (125, 8)       # equal batches, not 7 x 128 + 104
    assert plan(10, 128 * 128) == (10, 1)  # the reference's default render is one batch
    assert plan(1, N) == (1, 1) and plan(0, N) == (1, 0)
    # serial pipeline (B2PT_FLAG_NO_OVERLAP / one buffer set): as few batches as memory allows
    assert plan(128, N, sets=1) == (128, 1) and plan(300, N, sets=1) == (100, 3)
    # whole views as units: 225 views of 128 x 128 x 10 spp fit one batch; 64 views of 256^2 x 64 spp make four
    assert plan(225, 128 * 128 * 10) == (225, 1)
    assert plan(64, 256 * 256 * 64) == (16, 4)
    # a small target forces the split (what the tests use to exercise multi-batch paths)
    assert plan(6, 3072 * 6, max_paths_per_batch=2 * 3072 * 6 + 5) == (2, 3)
    assert plan(7, 100, max_paths_per_batch=50) == (1, 7)  # a unit larger than the target still gets its own batch
    # 4096^2: 8 samples per batch
    assert plan(512, 4096 * 4096) == (8, 64)
    # every unit is covered exactly once and no batch exceeds the target
    for units in (1, 3, 17, 255, 1024, 4097):
        for unit_paths in (1, 16384, N, 5 * N):
            for target in (N, 32 * Mi, 128 * Mi):
                per, nb = plan(units, unit_paths, target)
                assert per >= 1 and (nb - 1) * per < units <= nb * per
                assert per * unit_paths <= max(target, unit_paths)
    with pytest.raises(b2pt.B2ptError):
        plan(4, 0)
    with pytest.raises(b2pt.B2ptError):
        plan(4, N, sets=0)


@pytest.mark.parametrize("n", [1, 2, 7, 9, 300, 5000, 200000])
def test_bvh_builders_selfcheck(b2pt, n):
    """Binned-SAH tree + collapse into 8-wide compressed nodes on the many-sphere scene, validated on the host: every
    primitive reachable exactly once, child boxes nested, quantised boxes conservative with a step to spare."""
    bin_nodes, bin_depth, wide_nodes, wide_depth = b2pt.bvh_selfcheck(n)
    assert bin_nodes >= 1 and wide_nodes >= 1
    assert wide_depth <= max(1, (bin_depth + 1) // 2 + 1) and wide_depth <= 32
    assert wide_nodes <= max(1, bin_nodes // 2 + 1)
    # 8-wide: about n / 6 leaf-level nodes
    if n >= 300:
        assert wide_nodes < 0.45 * n


def test_fastdiv_recipe_is_exact():
    import numpy as np
    """csrc/b2pt_kernels.h make_fastdiv / b2pt_kernels.cu fastdiv (n / nPixels, n / sppPerView in the primary-ray index
    math): the Granlund-Montgomery round-up recipe, restated here, is exact for 32-bit n and every divisor in use."""
    def make(d):
        if d <= 1:
            return 0, None
        l = 0
        while (1 << l) < d:
            l += 1
        return ((1 << 32) * ((1 << l) - d)) // d + 1, l - 1

    def div(n, magic, shift):
        if shift is None:
            return n
        t = (n * magic) >> 32
        return ((t + ((n - t) >> 1)) & 0xffffffff) >> shift

    rng = np.random.default_rng(5)
    ds = [1, 2, 3, 5, 7, 10, 33 * 7, 128 * 128, 1024 * 1024, 1920 * 1080, 4096 * 4096, (1 << 26) - 1, 1 << 26, 1000, 1024,
          4095, 4097] + [int(x) for x in rng.integers(1, 1 << 26, 200)]
    for d in ds:
        magic, shift = make(d)
        assert magic < (1 << 32)
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, (1 << 32) - 1, (1 << 32) - d, (1 << 31), (1 << 31) - 1]
        ns += [int(x) for x in rng.integers(0, 1 << 32, 300)]
        ns += [k * d + r for k in (1, 7, ((1 << 32) - 1) // d) for r in (-1, 0, 1) if 0 <= k * d + r < (1 << 32)]
        for n in ns:
            if 0 <= n < (1 << 32):
                assert div(n, magic, shift) == n // d, (n, d)
