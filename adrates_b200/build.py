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


JAX_ADAPTER = os.path.join(HERE, "libadrates_b200_jax.so")


def jax_ffi_include():
    """Directory holding xla/ffi/api/ffi.h (shipped inside jaxlib), or None."""
    try:
        import jaxlib
    except ImportError:
        return None
    inc = os.path.join(os.path.dirname(jaxlib.__file__), "include")
    return inc if os.path.exists(os.path.join(inc, "xla", "ffi", "api", "ffi.h")) else None


def build_jax_ffi(force: bool = False):
    """The XLA FFI adapter (csrc/jax_ffi_adapter.cc -> libadrates_b200_jax.so, linked against libadrates_b200.so next to
    it).  Returns the path, or None where jaxlib's headers are not installed (this image)."""
    inc = jax_ffi_include()
    if inc is None:
        return None
    src = os.path.join(CSRC, "jax_ffi_adapter.cc")
    if not force and os.path.exists(JAX_ADAPTER) and os.path.getmtime(JAX_ADAPTER) > os.path.getmtime(src):
        return JAX_ADAPTER
    build()
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-std=c++17", "-O2", "-shared", "-Xcompiler", "-fPIC", "-I", inc, "-o", JAX_ADAPTER, src,
           "-L", HERE, "-ladrates_b200", "-Xlinker", "-rpath,$ORIGIN"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("jax ffi adapter build failed:\n" + res.stdout + res.stderr)
    return JAX_ADAPTER


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
