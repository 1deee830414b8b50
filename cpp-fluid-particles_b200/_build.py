"""Builds the native libraries in-tree (they travel to the GPU box with the snapshot):

  cpp-fluid-particles_b200/libsphk.so     nvcc, sm_100a: CUDA kernels + the C-ABI (include/sphk.h)
  cpp-fluid-particles_b200/libsphhost.so  g++: reference-shaped C++ classes + headless facade
                                          (include/sph_app.h), linked against libsphk.so
  cpp-fluid-particles_b200/sph_headless   g++: the reference application without its window (app/sph_headless.cpp),
                                          linked against the two libraries
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
FACADE = os.path.join(HERE, "facade")
INC = os.path.join(ROOT, "include")
NVCC = os.environ.get("SPHK_NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "/usr/bin/g++"
CUDA_INC = "/usr/local/cuda/include"
CUDA_LIB = "/usr/local/cuda/lib64"

LIBSPHK = os.path.join(HERE, "libsphk.so")
LIBHOST = os.path.join(HERE, "libsphhost.so")
APP = os.path.join(HERE, "app")
CLI = os.path.join(HERE, "sph_headless")

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-use_fast_math",
              "-Xcompiler", "-fPIC", "-I" + INC, "-I" + CSRC]


def _newer(target: str, sources: list[str]) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + "\n")
        raise RuntimeError("build failed: " + os.path.basename(cmd[-1]))


def build_sphk(force: bool = False, verbose_ptxas: bool = False) -> str:
    cus = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))
    deps = cus + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")] + [os.path.join(INC, "sphk.h")]
    if not force and _newer(LIBSPHK, deps):
        return LIBSPHK
    objs = []
    procs = []
    for cu in cus:
        obj = cu[:-3] + ".o"
        objs.append(obj)
        cmd = [NVCC] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose_ptxas else []) + ["-c", cu, "-o", obj]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + out + "\n")
            raise RuntimeError("nvcc failed")
        if verbose_ptxas:
            sys.stderr.write(out)
    _run([NVCC, "-shared", "-o", LIBSPHK] + objs + ["-lcudart", "-Xlinker", "-rpath=" + CUDA_LIB])
    return LIBSPHK


def build_host(force: bool = False) -> str:
    srcs = sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith(".cpp"))
    # the facade lives in its own directory so that its `#include "SPHSystem.h"` is resolved through -I
    # (this repo's headers here, the reference's in oracle/ref_build) and never through the source directory
    srcs += sorted(os.path.join(FACADE, f) for f in os.listdir(FACADE) if f.endswith(".cpp"))
    deps = srcs + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".h", ".hpp"))] + \
        [os.path.join(INC, "sphk.h"), os.path.join(INC, "sph_app.h"), LIBSPHK]
    if not force and _newer(LIBHOST, deps):
        return LIBHOST
    _run([CXX, "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-I" + INC, "-I" + HOST, "-I" + CUDA_INC, "-o", LIBHOST]
         + srcs + ["-L" + HERE, "-lsphk", "-L" + CUDA_LIB, "-lcudart", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + CUDA_LIB, "-Wl,-Bsymbolic"])
    return LIBHOST


def build_cli(force: bool = False) -> str:
    src = os.path.join(APP, "sph_headless.cpp")
    deps = [src, LIBHOST, LIBSPHK] + [os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".h", ".hpp"))]
    if not force and _newer(CLI, deps):
        return CLI
    _run([CXX, "-std=c++17", "-O2", "-Wall", "-I" + INC, "-I" + HOST, "-I" + CUDA_INC, "-o", CLI, src,
          "-L" + HERE, "-lsphhost", "-lsphk", "-L" + CUDA_LIB, "-lcudart", "-Wl,-rpath,$ORIGIN", "-Wl,-rpath," + CUDA_LIB])
    return CLI


def build_all(force: bool = False) -> None:
    build_sphk(force)
    build_host(force)
    build_cli(force)


if __name__ == "__main__":
    build_sphk(force="--force" in sys.argv, verbose_ptxas="-v" in sys.argv)
    if os.path.isdir(HOST) and any(f.endswith(".cpp") for f in os.listdir(HOST)):
        build_host(force="--force" in sys.argv)
    print("built", LIBSPHK)
