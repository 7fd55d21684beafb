"""Build libodf.so (the sm_100a CUDA kernels + C ABI) in-tree with nvcc.

Usage: python online-detection_b200/build_lib.py [--force]
The shared object lands next to this file (online-detection_b200/libodf.so); it is git-ignored but
travels to the GPU box with the working tree.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libodf.so")
SOURCES = ["odf_gauss_tile.cu", "odf_gauss_tile2.cu", "odf_panel.cu", "odf_panel16.cu", "odf_vec.cu", "odf_post.cu", "odf_precond.cu", "odf_rls.cu", "odf_tri.cu", "odf_api.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Wno-deprecated-gpu-targets"]


def _newer(a, b):
    return (not os.path.exists(b)) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [
        os.path.join(HERE, os.pardir, "include", "odf.h")]
    objs = []
    relink = force or not os.path.exists(LIB)
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or any(_newer(d, o) for d in deps):
            cmd = [NVCC] + ARCH + CFLAGS + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            subprocess.check_call(cmd)
            relink = True
    if relink:
        cmd = [NVCC, "-shared", "-Wno-deprecated-gpu-targets", "-o", LIB] + objs + [
            "-cudart", "static", "-lcublas", "-lcusolver"]
        if verbose:
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
