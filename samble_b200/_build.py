"""In-tree build of the native libraries (no JIT cache: the .so travels with the repo snapshot).

  libsamble_b200.so      nvcc, sm_100a only, CUDA runtime linked statically -> loads on a CPU-only box
  libsamble_hostcheck.so g++, host build of csrc/kalloc.h for the no-GPU tests
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT, "libsamble_b200.so")
HOSTLIB = os.path.join(OUT, "libsamble_hostcheck.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
NVCC_FLAGS = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps_mtime():
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "samble_b200.h")]
    return max(os.path.getmtime(f) for f in files)


def _run(cmd, log):
    p = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    if p.returncode != 0:
        raise RuntimeError(f"build failed: {' '.join(cmd)}\n{p.stdout}\n{p.stderr}")
    return p.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OUT, exist_ok=True)
    stamp = _deps_mtime()
    if not force and os.path.exists(LIB) and os.path.exists(HOSTLIB) and min(os.path.getmtime(LIB), os.path.getmtime(HOSTLIB)) >= stamp:
        return LIB
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT, src[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < stamp:
            err = _run([NVCC, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj], obj + ".log")
            if verbose:
                sys.stderr.write(err)
        return obj

    with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        objs = list(ex.map(compile_one, _sources()))
    _run([NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs], LIB + ".log")
    _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", HOSTLIB,
          os.path.join(CSRC, "hostcheck.cpp")], HOSTLIB + ".log")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
