"""In-tree nvcc build of the C-ABI CUDA library (sm_100a only).

``python -m avlen_b200._build`` (or ``__graft_entry__.build()``) compiles every
``csrc/*.cu`` into ``avlen_b200/libavlen_b200.so``.  nvcc cross-compiles
without a GPU.  The .so is git-ignored but travels with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ_DIR = os.path.join(HERE, "csrc", "_obj")
LIB_PATH = os.path.join(HERE, "libavlen_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON_FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC,-fvisibility=hidden",
                "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(path: str) -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(f.read())
    with open(path, "rb") as f:
        h.update(f.read())
    h.update(" ".join(ARCH_FLAGS + COMMON_FLAGS).encode())
    return h.hexdigest()


def _compile_one(src: str, verbose: bool) -> str:
    path = os.path.join(CSRC, src)
    obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
    stamp = obj + ".sha"
    dig = _digest(path)
    if os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read() == dig:
        return obj
    cmd = [NVCC, *ARCH_FLAGS, *COMMON_FLAGS, "-c", path, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{log}")
    if verbose:
        sys.stderr.write(log)
    with open(stamp, "w") as f:
        f.write(dig)
    return obj


PYHOST_SRC = os.path.join(CSRC, "py", "pyhost.c")
PYHOST_PATH = os.path.join(HERE, "_pyhost.so")


def build_pyhost(force: bool = False):
    """The CPython glue of batch_obs (csrc/py/pyhost.c -> avlen_b200/_pyhost.so), compiled with the host C compiler
    against this interpreter's headers.  Optional: without Python.h the pure-Python pointer loop is used instead."""
    import sysconfig
    inc = sysconfig.get_paths()["include"]
    if not os.path.exists(os.path.join(inc, "Python.h")):
        return None
    stamp = PYHOST_PATH + ".sha"
    with open(PYHOST_SRC, "rb") as f:
        dig = hashlib.sha256(f.read() + sys.version.encode()).hexdigest()
    if not force and os.path.exists(PYHOST_PATH) and os.path.exists(stamp) and open(stamp).read() == dig:
        return PYHOST_PATH
    cmd = [os.environ.get("CC", "gcc"), "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-I", inc, PYHOST_SRC, "-o", PYHOST_PATH]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("pyhost build failed:\n" + res.stdout + res.stderr)
    with open(stamp, "w") as f:
        f.write(dig)
    return PYHOST_PATH


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    build_pyhost(force)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(s, verbose), srcs))
    newest = max(os.path.getmtime(o) for o in objs)
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < newest:
        cmd = [NVCC, *ARCH_FLAGS, "-shared", "-o", LIB_PATH, *objs, "-lcudart", "-lcuda"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
