"""Compile csrc/*.cu for sm_100a into robot_ebert_b200/librebert_b200.so (in-tree, travels with gpurun).

nvcc cross-compiles without a GPU.  One object per source, compiled in parallel, then one shared library with
the static CUDA runtime (shares the primary context with torch, needs nothing but the driver at run time).
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "librebert_b200.so")
LIB_DEBUG = os.path.join(PKG, "librebert_b200_debug.so")      # -DREBERT_DEBUG: device-side bounds assertions (common.cuh)
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest(paths, extra=""):
    h = hashlib.sha256((" ".join(FLAGS) + extra).encode())
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(p.encode() + b"\0" + f.read())
    return h.hexdigest()


def _deps():
    hdr = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdr.append(os.path.join(os.path.dirname(PKG), "include", "rebert_b200.h"))
    return hdr


def build_native(force: bool = False, verbose: bool = False, debug: bool = False) -> str:
    """debug=True builds the assertion-carrying twin librebert_b200_debug.so (same sources, -DREBERT_DEBUG)."""
    obj_dir = OBJ + ("_debug" if debug else "")
    lib_path = LIB_DEBUG if debug else LIB
    flags = FLAGS + (["-DREBERT_DEBUG"] if debug else [])
    os.makedirs(obj_dir, exist_ok=True)
    srcs, hdrs = _sources(), _deps()
    stamp = os.path.join(obj_dir, "stamp")
    want = _digest(srcs + hdrs, "debug" if debug else "")
    if not force and os.path.exists(lib_path) and os.path.exists(stamp) and open(stamp).read() == want:
        return lib_path
    if not os.path.exists(NVCC):
        raise RuntimeError(f"nvcc not found at {NVCC} and {lib_path} is missing or stale")

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        key = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".key")
        k = _digest([src] + hdrs, "debug" if debug else "")
        if not force and os.path.exists(obj) and os.path.exists(key) and open(key).read() == k:
            return obj, ""
        cmd = [NVCC, *flags, "-Xptxas", "-v", "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        with open(key, "w") as f:
            f.write(k)
        with open(os.path.join(obj_dir, os.path.basename(src)[:-3] + ".ptxas.log"), "w") as f:
            f.write(r.stderr)
        return obj, r.stderr

    with cf.ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
        results = list(ex.map(compile_one, srcs))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            sys.stderr.write(log)
    cmd = [NVCC, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static",
           "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return lib_path


def build_pycall(force: bool = False) -> str:
    """Compile csrc/pycall.c (CPython side door into rebert_recommend_host: saves the ctypes call overhead on the request
    path) with the host C compiler into robot_ebert_b200/_pycall<EXT_SUFFIX>.  Optional: without it the same call goes
    through ctypes."""
    import sysconfig
    src = os.path.join(CSRC, "pycall.c")
    out = os.path.join(PKG, "_pycall" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
    stamp = os.path.join(OBJ, "pycall.stamp")
    os.makedirs(OBJ, exist_ok=True)
    want = _digest([src], sysconfig.get_paths()["include"])
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read() == want:
        return out
    cc = os.environ.get("CC", "gcc")
    cmd = [cc, "-O2", "-std=c11", "-Wall", "-Wextra", "-shared", "-fPIC", "-I", sysconfig.get_paths()["include"], src, "-o", out]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"{cc} failed for {src}:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want)
    return out


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose=True, debug="--debug" in sys.argv))
    print(build_pycall(force="--force" in sys.argv))
