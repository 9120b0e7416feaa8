"""Build the CUDA shared library in-tree (adrates_b200/libadrates_b200.so) for sm_100a."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = [os.path.join(CSRC, "cav_api.cu"), os.path.join(CSRC, "cav_book.cu"), os.path.join(CSRC, "cav_comm.cu")]
DEPS = SOURCES + [os.path.join(CSRC, h) for h in ("cav_kernels.cuh", "cav_ctx.h", "cav_book_core.h")] + \
    [os.path.join(HERE, "..", "include", "adrates_b200.h")]
LIB = os.path.join(HERE, "libadrates_b200.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC,-fopenmp"]


def needs_build() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    extra = os.environ.get("CAV_NVCC_EXTRA", "").split()          # experiment switches (-DNAME=value)
    objs = [os.path.join(HERE, os.path.basename(s)[:-3] + ".o") for s in SOURCES]

    def compile_one(pair):
        src, obj = pair
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        return subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(len(SOURCES)) as ex:       # one nvcc per translation unit, side by side
        results = list(ex.map(compile_one, zip(SOURCES, objs)))
    for res in results:
        if res.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
        if verbose:
            sys.stderr.write(res.stderr)
    res = subprocess.run([nvcc, "-shared", "-Xcompiler", "-fPIC,-fopenmp", "-gencode", "arch=compute_100a,code=sm_100a",
                          "-o", LIB] + objs, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + res.stdout + res.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
