"""In-tree build of the native parts (no JIT cache: the .so files travel with the repo snapshot).

  sapien_b200/csrc/libss_b200.so            CUDA kernels + engine + C ABI      (nvcc, sm_100a)
  sapien_b200/_simsense_b200.<abi>.so       pybind11 module over the C ABI     (g++)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB = os.path.join(CSRC, "libss_b200.so")
EXT = os.path.join(ROOT, "_simsense_b200" + sysconfig.get_config_var("EXT_SUFFIX"))
CU_SOURCES = ["engine.cu", "front.cu", "cost.cu", "aggr.cu", "generic.cu", "post.cu"]
HEADERS = ["common.cuh", "kernels.h", os.path.join("..", "..", "include", "ss_b200.h")]
NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--threads", "0"]


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)


def build_lib(force: bool = False, verbose: bool = False) -> str:
    srcs = [os.path.join(CSRC, s) for s in CU_SOURCES]
    deps = srcs + [os.path.normpath(os.path.join(CSRC, h)) for h in HEADERS]
    if force or _stale(LIB, deps):
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        if not os.path.exists(nvcc):
            raise RuntimeError("nvcc not found and no prebuilt libss_b200.so")
        from concurrent.futures import ThreadPoolExecutor

        objs, jobs = [], []
        for s in CU_SOURCES:  # separate objects: only changed sources recompile, and they compile side by side
            o = os.path.join(CSRC, s[:-3] + ".o")
            if force or _stale(o, [os.path.join(CSRC, s)] + deps[len(srcs):]):
                if verbose:
                    print("nvcc", s, file=sys.stderr)
                jobs.append([nvcc, *NVCC_FLAGS, "-c", s, "-o", o])
            objs.append(o)
        with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as pool:
            list(pool.map(_run, jobs))
        _run([nvcc, "-shared", "-o", LIB, *objs])
    return LIB


def build_ext(force: bool = False) -> str:
    src = os.path.join(CSRC, "pybind.cpp")
    deps = [src, os.path.normpath(os.path.join(CSRC, HEADERS[2])), LIB]
    if force or _stale(EXT, deps):
        import pybind11

        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        _run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden",
              "-I" + pybind11.get_include(), "-I" + sysconfig.get_paths()["include"],
              src, "-o", EXT, "-L" + CSRC, "-lss_b200", "-Wl,-rpath,$ORIGIN/csrc"])
    return EXT


def build_all(force: bool = False, verbose: bool = False) -> None:
    build_lib(force, verbose)
    build_ext(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose=True)
    print(LIB)
    print(EXT)
