"""Builds the native parts in-tree: libsfgpu.so (CUDA, sm_100a only) and the C host library.

nvcc cross-compiles without a GPU; the built .so files travel to the GPU box with the snapshot
(they are git-ignored, not gpurun-ignored).
"""
from __future__ import annotations

import os
import shutil
import subprocess

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
HOST = os.path.join(PKG, "host")
LIB_GPU = os.path.join(PKG, "libsfgpu.so")
LIB_HOST = os.path.join(PKG, "libsfhost.so")
CLI = os.path.join(PKG, "sigfish-b200")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # B200 only, no PTX for other archs
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",            # the reference is built without FMA contraction
    "-prec-div=true", "-prec-sqrt=true",
    "-Xcompiler", "-fPIC",
]
OBJ_DIR = os.path.join(ROOT, "build", "obj")


def _newer(target: str, sources) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found: libsfgpu.so cannot be built (there is no CPU fallback)")


def gpu_sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + \
        [os.path.join(ROOT, "include", "sfgpu.h")]


def build_gpu(force: bool = False, verbose: bool = False) -> str:
    """libsfgpu.so = sfgpu.cu (C-ABI + most kernels) + the sf_inst_*.cu translation units that hold the ~250
    instantiations of the pair DTW kernel; the translation units are compiled side by side."""
    from concurrent.futures import ThreadPoolExecutor

    srcs = gpu_sources()
    if not force and _newer(LIB_GPU, srcs):
        return LIB_GPU
    os.makedirs(OBJ_DIR, exist_ok=True)
    units = [s for s in srcs if s.endswith(".cu")]
    objs = [os.path.join(OBJ_DIR, os.path.basename(u)[:-3] + ".o") for u in units]

    def compile_unit(uo):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas=-v"] if verbose else []) + ["-c", "-o", uo[1], uo[0]]
        subprocess.run(cmd, check=True)

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 1)) as ex:
        list(ex.map(compile_unit, zip(units, objs)))
    subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB_GPU] + objs, check=True)
    return LIB_GPU


def host_sources():
    if not os.path.isdir(HOST):
        return []
    return sorted(os.path.join(HOST, f) for f in os.listdir(HOST) if f.endswith((".c", ".h")))


def build_host(force: bool = False) -> str | None:
    srcs = host_sources()
    c_files = [s for s in srcs if s.endswith(".c")]
    if not c_files:
        return None
    deps = srcs + [os.path.join(ROOT, "include", "sfgpu.h")]
    lib_c = [s for s in c_files if os.path.basename(s) != "main.c"]
    if force or not _newer(LIB_HOST, deps):
        subprocess.run(["gcc", "-O2", "-g", "-std=c99", "-Wall", "-fPIC", "-shared", "-D_GNU_SOURCE",
                        "-I", os.path.join(ROOT, "include"), "-o", LIB_HOST] + lib_c +
                       ["-L", PKG, "-lsfgpu", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread", "-lm"], check=True)
    main_c = os.path.join(HOST, "main.c")
    if os.path.exists(main_c) and (force or not _newer(CLI, deps)):
        subprocess.run(["gcc", "-O2", "-g", "-std=c99", "-Wall", "-D_GNU_SOURCE",
                        "-I", os.path.join(ROOT, "include"), "-o", CLI, main_c,
                        "-L", PKG, "-lsfhost", "-lsfgpu", "-Wl,-rpath,$ORIGIN", "-lz", "-lpthread", "-lm"], check=True)
    return LIB_HOST


def build_all(force: bool = False) -> None:
    build_gpu(force)
    build_host(force)


if __name__ == "__main__":
    import sys
    build_all(force="--force" in sys.argv)
    print(LIB_GPU)
