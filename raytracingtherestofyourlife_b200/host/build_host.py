"""Builds the C++ facade (libb2pt_facade.so) and the CornellBox_b2pt driver against libb2pt.so, in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
CXX = "/usr/bin/g++"  # $CXX in this image lacks its spec files; use the system compiler
FLAGS = ["-std=c++17", "-O2", "-fPIC", "-ffp-contract=off", "-Wall", "-Wextra", "-Wno-unused-parameter",
         "-I", os.path.join(HERE, "vtkm_shim"), "-I", HERE]
FACADE_SRC = ["b2pt_facade.cxx", "pathtracing/Camera.cxx", "MapperPathTracer.cxx", "CornellBox.cpp"]
LIB = os.path.join(HERE, "libb2pt_facade.so")
EXE = os.path.join(HERE, "CornellBox_b2pt")
TEST = os.path.join(HERE, "test_facade")


def _newer(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _all_sources():
    out = []
    for d, _, files in os.walk(HERE):
        out += [os.path.join(d, f) for f in files if f.endswith((".h", ".cxx", ".cpp", ".cc", ".py"))]
    out.append(os.path.join(PKG, "..", "include", "b2pt.h"))
    return out


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("host build failed")


def build_host(force=False):
    deps = _all_sources() + [os.path.join(PKG, "libb2pt.so")]
    link = ["-L", PKG, "-lb2pt", "-Wl,-rpath,$ORIGIN/..", "-Wl,-rpath,$ORIGIN"]
    if force or _newer(LIB, deps):
        _run([CXX] + FLAGS + ["-shared", "-o", LIB] + [os.path.join(HERE, s) for s in FACADE_SRC] + link)
    if force or _newer(EXE, deps + [LIB]):
        _run([CXX] + FLAGS + ["-o", EXE, os.path.join(HERE, "main.cc"), "-L", HERE, "-lb2pt_facade"] + link)
    if os.path.exists(os.path.join(HERE, "test_facade.cc")) and (force or _newer(TEST, deps + [LIB])):
        _run([CXX] + FLAGS + ["-o", TEST, os.path.join(HERE, "test_facade.cc"), "-L", HERE, "-lb2pt_facade"] + link)
    return LIB, EXE


if __name__ == "__main__":
    print(build_host(force="--force" in sys.argv))
