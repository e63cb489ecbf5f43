"""Builds libvt_b200.so (hand-written sm_100a CUDA + the C ABI of include/vt_b200.h) in-tree."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libvt_b200.so")
SOURCES = ["vt_api.cu", "full_step.cu", "full_step_async.cu", "full_step_tma.cu", "poisson.cu", "halo.cu", "tucker.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-O3", "--shared", "-cudart", "shared",
]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "vt_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    cmd = [nvcc()] + NVCC_FLAGS + list(extra) + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
        print(" ".join(cmd))
    # use the system host compiler; the image exports CXX/CC pointing at a wrapper nvcc cannot use
    env = dict(os.environ)
    cmd += ["-ccbin", "/usr/bin/g++"]
    subprocess.check_call(cmd, env=env)
    return LIB


if __name__ == "__main__":
    build_lib(force=True, verbose="-v" in sys.argv)
    print(LIB)
