"""The oracle is test infrastructure: nothing in the product may import, link or execute it."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "raytracingtherestofyourlife_b200")


def product_sources():
    for d, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".cxx", ".cc")):
                yield os.path.join(d, f)
    yield os.path.join(ROOT, "include", "b2pt.h")


def test_product_never_references_the_oracle():
    pat = re.compile(r"(import\s+oracle|from\s+oracle|b2pt_oracle|orc_[a-z_]+\s*\()")
    for path in product_sources():
        assert not pat.search(open(path).read()), path


def test_library_does_not_link_the_oracle(b2pt):
    out = subprocess.check_output(["ldd", b2pt.LIB_PATH], text=True)
    assert "oracle" not in out
    syms = subprocess.check_output(["nm", "-D", b2pt.LIB_PATH], text=True)
    assert "orc_" not in syms


def test_nothing_reads_the_reference_tree_at_run_time():
    for path in list(product_sources()) + [os.path.join(ROOT, "bench.py"), os.path.join(ROOT, "__graft_entry__.py")]:
        if not os.path.exists(path):
            continue
        src = open(path).read()
        src = re.sub(r"#.*|//.*", "", src)
        assert "/root/reference" not in src, path
