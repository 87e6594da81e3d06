"""In-tree build of the engine's shared libraries with nvcc for sm_100a.

  beast-mcmc_b200/csrc/libhmsbeagle.so      the C ABI (include/libhmsbeagle_b200.h), CUDA runtime linked statically
  beast-mcmc_b200/csrc/libhmsbeagle-jni.so  the JNI shim BEAST's lib/beagle.jar binds (System.loadLibrary("hmsbeagle-jni"))
  oracle/liboracle_cpu.so                   the CPU restatement used as checker / cpu_baseline (test infrastructure)

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
NVCC = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd, verbose):
    if verbose:
        print(" ".join(cmd), flush=True)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))
    if verbose and (r.stdout or r.stderr):
        print(r.stdout + r.stderr)


def lib_path(name="libhmsbeagle.so"):
    return os.path.join(CSRC, name)


def build_engine(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in ("api.cu", "kernels.cu", "patterns.cu")]
    deps = srcs + [os.path.join(CSRC, "engine.h"), os.path.join(ROOT, "include", "libhmsbeagle_b200.h")]
    out = lib_path()
    if force or _stale(out, deps):
        _run([NVCC, *ARCH, "-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-fvisibility=hidden",
              "-shared", "-cudart", "static", "-Xptxas", "-v" if verbose else "-O3",
              "-o", out, *srcs], verbose)
    return out


def build_jni(force=False, verbose=False):
    src = os.path.join(CSRC, "jni_shim.cpp")
    if not os.path.exists(src):
        return None
    out = lib_path("libhmsbeagle-jni.so")
    deps = [src, os.path.join(ROOT, "include", "jni_min.h"), os.path.join(ROOT, "include", "libhmsbeagle_b200.h")]
    if force or _stale(out, deps):
        _run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-I", os.path.join(ROOT, "include"),
              "-o", out, src, "-L", CSRC, "-lhmsbeagle", "-Wl,-rpath,$ORIGIN"], verbose)
    return out


def build_oracle(force=False, verbose=False):
    src = os.path.join(ROOT, "oracle", "beagle_cpu.c")
    if not os.path.exists(src):
        return None
    out = os.path.join(ROOT, "oracle", "liboracle_cpu.so")
    if force or _stale(out, [src]):
        _run(["gcc", "-O3", "-march=native", "-std=gnu11", "-fPIC", "-shared", "-pthread", "-o", out, src, "-lm"],
             verbose)
    return out


def build_all(force=False, verbose=False):
    return build_engine(force, verbose), build_jni(force, verbose), build_oracle(force, verbose)


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv, verbose=True))
