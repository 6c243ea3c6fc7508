"""Build art_b200/libart_hotpath.so (the C-ABI library) in-tree with nvcc for sm_100a.

No torch involved: the library is plain CUDA runtime + the C-ABI in include/art_hotpath.h.
-fmad=false: the reference's arithmetic is uncontracted SSE2; FMA contraction would change results.
"""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libart_hotpath.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-pthread",
    "-ccbin", "/usr/bin/g++",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "art_hotpath.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return LIB
    from concurrent.futures import ThreadPoolExecutor
    hdrs = glob.glob(os.path.join(CSRC, "*.h")) + glob.glob(os.path.join(CSRC, "*.cuh")) + \
        [os.path.join(HERE, "..", "include", "art_hotpath.h"), __file__]
    objs, jobs = [], []
    for src in sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        # per-object staleness: a translation unit is recompiled when it or any header is newer than its object
        if force or verbose or not os.path.exists(obj) or any(os.path.getmtime(d) > os.path.getmtime(obj) for d in [src] + hdrs):
            jobs.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj])
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(subprocess.check_call, jobs))
    subprocess.check_call([NVCC, "-shared", "-o", LIB] + objs + ["-lcudart", "-ldl", "-ccbin", "/usr/bin/g++"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
