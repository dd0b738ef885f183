"""The C-ABI library loads and exports every symbol include/b2pt.h declares (no compute without a GPU)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "b2pt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2pt_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    names = declared_functions()
    for must in ("b2pt_create", "b2pt_destroy", "b2pt_set_scene", "b2pt_build_bvh", "b2pt_set_camera", "b2pt_seed",
                 "b2pt_render", "b2pt_primary_hits", "b2pt_allreduce", "b2pt_read_color", "b2pt_get_stats"):
        assert must in names  # SURVEY.md 8b proposed C-ABI


def test_library_exports_every_declared_symbol(b2pt):
    L = ctypes.CDLL(b2pt.LIB_PATH)
    for name in declared_functions():
        assert hasattr(L, name), "libb2pt.so does not export %s" % name


def test_binding_covers_every_declared_symbol(b2pt):
    assert sorted(b2pt.SIGNATURES) == declared_functions()


def test_no_unexpected_exports(b2pt):
    out = subprocess.check_output(["nm", "-D", "--defined-only", b2pt.LIB_PATH], text=True)
    exported = sorted(l.split()[-1] for l in out.splitlines() if " T b2pt_" in l)
    assert exported == declared_functions()


def test_library_is_sm100a_only(b2pt):
    out = subprocess.run(["cuobjdump", "-lelf", b2pt.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_version_and_error_string(b2pt):
    assert b2pt.lib().b2pt_version() >= 100
    assert isinstance(b2pt.lib().b2pt_last_error(), bytes)


def test_fails_loudly_without_a_gpu(b2pt):
    """No CPU fallback: creating a context on a box without a B200 must raise, not degrade."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(b2pt.B2ptError) as e:
        b2pt.Context(0)
    assert e.value.code == b2pt.ERR_CUDA
    assert "no CPU fallback" in str(e.value)
