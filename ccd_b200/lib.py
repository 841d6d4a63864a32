"""Build and load the C-ABI shared library (include/ccd_b200.h) with ctypes.

The library is compiled in-tree with plain nvcc for sm_100a (no torch headers in the ABI) so the built .so
travels with the repository snapshot.  There is NO fallback: every op raises if the library is missing.
"""
import ctypes
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SO_PATH = os.environ.get("CCD_LIB", os.path.join(HERE, "libccd_b200.so"))   # CCD_LIB: experiment builds (tools/)
HEADER = os.path.join(os.path.dirname(HERE), "include", "ccd_b200.h")
SOURCES = ["gemm_umma.cu", "mhsa_fwd.cu", "mhsa_bwd.cu", "rowwise.cu", "dino_loss.cu", "charseg.cu", "seghead.cu", "optim.cu", "decoder.cu", "abi.cu"]

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "--use_fast_math",
              "-Xcompiler", "-fPIC", "-cudart", "static"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def _stale():
    if not os.path.exists(SO_PATH):
        return True
    t = os.path.getmtime(SO_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [HEADER]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, defines=(), out=None, tag=""):
    """Compile every CUDA source for sm_100a into ccd_b200/libccd_b200.so (cross-compiles without a GPU).
    `defines` / `out` / `tag`: experiment builds with extra -D switches into another .so (tools/build_variants.py)."""
    out = out or SO_PATH
    if not force and not defines and not _stale():
        return out
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        obj = os.path.join(HERE, "build", src.replace(".cu", tag + ".o"))
        cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}" for d in defines], "-I", CSRC, "-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        log, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose and log.strip():
            print(log)
    cmd = [_nvcc(), "-shared", "-o", out, *objs, "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return out


def declared_symbols():
    """Every `int ccd_*(` entry point declared in include/ccd_b200.h."""
    with open(HEADER) as f:
        return sorted(set(re.findall(r"^int\s+(ccd_\w+)\s*\(", f.read(), flags=re.M)))


_LIB = None


def load(build_if_missing=True):
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(SO_PATH):
        if not build_if_missing:
            raise RuntimeError(f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(ccd_b200 has no CPU / PyTorch fallback)")
        build()
    lib = ctypes.CDLL(SO_PATH)
    for name in declared_symbols():
        fn = getattr(lib, name)          # AttributeError here = header/library mismatch: fail loudly
        fn.restype = ctypes.c_int
    _LIB = lib
    return lib


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
    print(SO_PATH, "exports", len(declared_symbols()), "entry points")
