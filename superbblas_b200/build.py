"""Build libsuperbblas_b200.so (CUDA kernels + host planner + C ABI) for sm_100a, in tree.

    python -m superbblas_b200.build [--force]

nvcc cross-compiles without a GPU.  The .so lands in superbblas_b200/lib/ (git-ignored; it travels
to the GPU box with the working tree).
"""
import concurrent.futures
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "lib")
LIB = os.path.join(OUT, "libsuperbblas_b200.so")
SOURCES = ["geometry.cpp", "plan.cpp", "contract_plan.cpp", "runtime.cpp", "capi.cpp",
           "kernels_copy.cu", "kernels_contract.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
         "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++"]


def _stamp(path):
    h = hashlib.sha1()
    for f in sorted(os.listdir(SRC)) + ["../../include/superbblas_b200.h"]:
        with open(os.path.join(SRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src):
    obj = os.path.join(OUT, src + ".o")
    cmd = [NVCC] + FLAGS + ["-c", os.path.join(SRC, src), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    return obj


def build(force=False, verbose=True):
    os.makedirs(OUT, exist_ok=True)
    stamp_file = os.path.join(OUT, "stamp")
    stamp = _stamp(SRC)
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and \
            open(stamp_file).read() == stamp:
        return LIB
    if verbose:
        print("[superbblas_b200] compiling for sm_100a ...", file=sys.stderr)
    with concurrent.futures.ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(_compile, SOURCES))
    cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-ccbin", "/usr/bin/g++", "-lpthread", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
