"""Builds libnrvit.so (the C-ABI library of sm_100a kernels) in-tree with nvcc.

Usage: python build.py [--force] [--verbose]
The output lands in noise-robust-vit_b200/lib/ so that it travels to the GPU box with the
repository snapshot (the .so is git-ignored, not gpurun-ignored).
"""
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libnrvit.so")
STAMP = os.path.join(LIBDIR, "libnrvit.stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",  # NOT -arch=sm_100a: that also emits plain compute_100 PTX
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def _nccl_flags():
    """Compile against the NCCL that torch loads (pip wheel, 2.28.x); link by SONAME only."""
    try:
        import nvidia.nccl as n  # type: ignore
        root = os.path.dirname(n.__file__) if getattr(n, "__file__", None) else list(n.__path__)[0]
        inc, lib = os.path.join(root, "include"), os.path.join(root, "lib")
        if os.path.exists(os.path.join(inc, "nccl.h")):
            return ["-I" + inc, "-L" + lib, "-l:libnccl.so.2", "-Xlinker", "-rpath," + lib]
    except Exception:
        pass
    if os.path.exists("/usr/include/nccl.h"):
        return ["-l:libnccl.so.2"]
    return None


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    os.makedirs(LIBDIR, exist_ok=True)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP):
        with open(STAMP) as fh:
            if fh.read().strip() == digest:
                return LIB
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        if os.path.exists(LIB):
            return LIB  # GPU box without a toolchain change: use the prebuilt library
        raise RuntimeError("nvcc not found and no prebuilt libnrvit.so")
    srcs = sources()
    link = []
    if any(s.endswith("comm.cu") for s in srcs):
        nccl = _nccl_flags()
        if nccl is None:
            srcs = [s for s in srcs if not s.endswith("comm.cu")]
        else:
            link = [f for f in nccl if f.startswith("-I")] + ["-ldl"]   # nccl.h only: comm.cu binds libnccl with dlopen
    # one nvcc -c per translation unit, in parallel (the tcgen05 kernels take 20-60 s each), then one link
    objdir = os.path.join(LIBDIR, "obj")
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    inc = [f for f in link if f.startswith("-I")]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + compile_flags + inc + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
        res = subprocess.run(cmd, capture_output=True, text=True)
        return src, obj, res

    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=min(len(srcs), os.cpu_count() or 4)) as ex:
        results = list(ex.map(compile_one, srcs))
    failed = False
    for src, obj, res in results:
        if verbose or res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
        failed = failed or res.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc] + NVCC_FLAGS + [f for f in link if not f.startswith("-I")] + ["-o", LIB] + [obj for _, obj, _ in results]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed (exit %d)" % res.returncode)
    with open(STAMP, "w") as fh:
        fh.write(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
