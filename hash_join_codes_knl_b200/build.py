"""Builds libhjb200.so (the CUDA hot path + C ABI) in-tree with nvcc for sm_100a, and the
C++ host programs npj / phj / cpra / write that drive it (bin/)."""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(HERE, "libhjb200.so")
BIN = os.path.join(HERE, "bin")
KERNEL_SOURCES = ["radix.cu", "part_join.cu", "npj.cu", "skew.cu", "stage.cu", "gen.cu", "capi.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall"] + os.environ.get("HJB_NVCC_EXTRA", "").split()


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libhjb200.so cannot be built (there is no CPU fallback)")
    return nvcc


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def build_library(force=False, verbose=False):
    srcs = [os.path.join(CSRC, f) for f in KERNEL_SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in ("hj_device.cuh", "hj_internal.h")] + \
        [os.path.join(os.path.dirname(HERE), "include", "hjb200.h")]
    if force or _stale(LIB, deps):
        cmd = [_nvcc(), "-shared"] + NVCC_FLAGS + srcs + ["-o", LIB]
        if verbose:
            print(" ".join(cmd))
        subprocess.check_call(cmd)
    return LIB


def build_programs(force=False, verbose=False):
    """The reference's CLI (./npj|./phj|./cpra [#threads] [outer] [inner], ./write ...)."""
    os.makedirs(BIN, exist_ok=True)
    common = [os.path.join(HOST, "hj_host.h"), LIB]
    out = []
    for name in ("npj", "phj", "cpra", "write"):
        src = os.path.join(HOST, name + ".cpp")
        if not os.path.exists(src):
            continue
        exe = os.path.join(BIN, name)
        if force or _stale(exe, [src] + common):
            cmd = [_nvcc(), "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", os.path.join(os.path.dirname(HERE), "include"), src,
                   "-o", exe, "-L", HERE, "-lhjb200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/..", "-lpthread"]
            if verbose:
                print(" ".join(cmd))
            subprocess.check_call(cmd)
        out.append(exe)
    return out


if __name__ == "__main__":
    build_library(force="--force" in sys.argv, verbose=True)
    build_programs(force="--force" in sys.argv, verbose=True)
