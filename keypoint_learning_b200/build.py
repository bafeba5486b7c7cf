"""In-tree build of libkpl_b200.so (CUDA kernels + C ABI) and the C++ host tools, sm_100a only.

nvcc cross-compiles without a GPU.  -fmad=false is part of the arithmetic contract of the hot path
(see csrc/kpl_math.cuh), not an optimisation switch.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libkpl_b200.so")
TEST_DETECTOR = os.path.join(HERE, "TestDetector")
PCD_TOOL = os.path.join(HERE, "pcd_tool")

CU_SOURCES = ["capi.cu", "grid.cu", "normals.cu", "features.cu", "forest.cu", "nms.cu", "shard.cu", "organized.cu", "forest_yaml.cpp"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2", "--expt-relaxed-constexpr", "-cudart", "static",
]


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _host_cxx() -> str:
    # the image exports CXX=/opt/gcc/bin/g++ (a wrapper); the distro compiler is the reliable one
    return "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else (shutil.which("g++") or "g++")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.join(CSRC, h) for h in ("kpl_internal.h", "kpl_math.cuh", "forest.cuh")] + [os.path.join(ROOT, "include", "kpl.h"), __file__]
    if not force and not _stale(LIB, deps):
        return LIB
    objdir = os.path.join(HERE, "build")
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    extra = os.environ.get("KPL_NVCC_EXTRA", "").split()      # tuning experiments only, e.g. -DKPL_TREES_IN_FLIGHT=8
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-ccbin", _host_cxx(), "-x", "cu", "-c", s, "-o", o]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    link = [_nvcc(), "-shared", "-o", LIB, *objs, "-ccbin", _host_cxx(), "-gencode", "arch=compute_100a,code=sm_100a",
            "-cudart", "static", "-lz", "-ldl", "-Xlinker", "--no-undefined"]
    subprocess.check_call(link)
    return LIB


def build_host(force: bool = False) -> str:
    """C++ facade + TestDetector CLI (links against libkpl_b200.so with an $ORIGIN rpath)."""
    if not os.path.isdir(HOST):
        return ""
    deps = [os.path.join(HOST, f) for f in os.listdir(HOST)] + [LIB, os.path.join(ROOT, "include", "kpl.h")]
    common = [os.path.join(HOST, f) for f in sorted(os.listdir(HOST)) if f.endswith(".cpp") and not f.startswith("main_")]
    for main, exe in (("main_test_detector.cpp", TEST_DETECTOR), ("main_pcd_tool.cpp", PCD_TOOL)):
        if not os.path.exists(os.path.join(HOST, main)) or (not force and not _stale(exe, deps)):
            continue
        cmd = [_host_cxx(), "-O2", "-std=c++17", "-Wall", "-Wextra", "-I", os.path.join(ROOT, "include"), "-I", HOST,
               os.path.join(HOST, main), *common, "-o", exe, "-L", HERE, "-lkpl_b200", "-lpthread", "-Wl,-rpath,$ORIGIN"]
        subprocess.check_call(cmd)
    return TEST_DETECTOR


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_host(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(LIB)
