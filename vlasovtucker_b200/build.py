"""Builds libvt_b200.so (hand-written sm_100a CUDA + the C ABI of include/vt_b200.h) in-tree.

One nvcc process per translation unit, run concurrently, then one link step: the objects live in
vlasovtucker_b200/build/obj (git-ignored like the library itself)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.path.join(HERE, "build", "obj")
LIB = os.path.join(LIBDIR, "libvt_b200.so")
# (source, object, extra flags): the general Tucker kernel is compiled once per instantiation (three objects from tucker_inst.cu)
SOURCES = ["vt_api.cu", "full_step.cu", "full_step_tma.cu", "poisson.cu", "halo.cu", "tucker.cu", "tucker_slab.cu", "group.cu",
           ("tucker_inst.cu", "tucker_inst64.o", ["-DVT_TUCKER_NM=64"]), ("tucker_inst.cu", "tucker_inst32.o", ["-DVT_TUCKER_NM=32"]),
           ("tucker_inst.cu", "tucker_inst16.o", ["-DVT_TUCKER_NM=16"])]
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMPILE_FLAGS = ARCH_FLAGS + ["-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]
# the system host compiler; the image exports CXX/CC pointing at a wrapper nvcc cannot use
HOST_CXX = ["-ccbin", "/usr/bin/g++"]


def nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "vt_b200.h")]


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_lib(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(OBJDIR, exist_ok=True)
    sources = [s if isinstance(s, tuple) else (s, s[:-3] + ".o", []) for s in SOURCES]
    sources = [s for s in sources if os.path.exists(os.path.join(CSRC, s[0]))]
    headers_t = max(os.path.getmtime(d) for d in _deps() if d.endswith((".h", ".inl")))

    def compile_one(item):
        src, objname, flags = item
        obj = os.path.join(OBJDIR, objname)
        path = os.path.join(CSRC, src)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(path), headers_t):
            return obj, ""
        cmd = [nvcc()] + COMPILE_FLAGS + list(flags) + list(extra) + (["-Xptxas", "-v"] if verbose else []) + HOST_CXX + ["-c", path, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src} {flags}:\n{r.stdout}\n{r.stderr}")
        return obj, r.stdout + r.stderr

    with ThreadPoolExecutor(max_workers=min(len(sources), os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources))
    for _, log in results:
        if log.strip():
            print(log.rstrip(), file=sys.stderr if not verbose else sys.stdout)
    link = [nvcc()] + ARCH_FLAGS + ["--shared", "-cudart", "shared"] + HOST_CXX + ["-o", LIB] + [o for o, _ in results]
    subprocess.check_call(link)
    return LIB


if __name__ == "__main__":
    build_lib(force=True, verbose="-v" in sys.argv)
    print(LIB)
