"""Build recipe for libb2pt.so (hand-written sm_100a kernels + C-ABI), in-tree.

nvcc cross-compiles without a GPU.  Flags:
  -gencode arch=compute_100a,code=sm_100a   Blackwell B200 only
  -fmad=false                                no FMA contraction: geometric predicates stay bit-identical
                                             to the reference's x86-64 (no-FMA) arithmetic
  -lineinfo                                  ncu source page maps SASS to these files
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("B2PT_LIB") or os.path.join(HERE, "libb2pt.so")  # B2PT_LIB: experiment builds
SOURCES = ["b2pt_kernels.cu", "b2pt_api.cu", "b2pt_lbvh.cu", "b2pt_scene.cpp"]
HEADERS = ["b2pt_device.cuh", "b2pt_kernels.h", "b2pt_types.h", "b2pt_bvh.h", "b2pt_wide.h", "b2pt_lbvh.h", "../../include/b2pt.h"]
NVCC = os.environ.get("B2PT_NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"  # $CXX in this image points at a g++ without OpenMP/specs; use the system one

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-fmad=false",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-fopenmp",
]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    """Compile libb2pt.so if sources are newer than the library. Returns its path."""
    if os.environ.get("B2PT_LIB"):
        return LIB
    if not force and not _stale():
        return LIB
    def compile_one(src):
        obj = os.path.join(CSRC, os.path.splitext(src)[0] + ".o")
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        if src.endswith(".cpp"):
            cmd = [NVCC] + NVCC_FLAGS + ["-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, r

    from concurrent.futures import ThreadPoolExecutor
    objs = []
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as pool:  # the translation units compile side by side
        for src, obj, r in pool.map(compile_one, SOURCES):
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed for %s" % src)
            objs.append(obj)
    cmd = [NVCC, "-shared", "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fopenmp",
           "-o", LIB] + objs
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose="-v" in sys.argv))
