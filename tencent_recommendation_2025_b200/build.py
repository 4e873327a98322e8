"""Build libtgr_embed.so (the C-ABI CUDA library) in-tree for sm_100a with nvcc.

    python -m tencent_recommendation_2025_b200.build [--force]

nvcc cross-compiles without a GPU. Objects are cached under csrc/build/ keyed on source mtimes.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(ROOT, "include")
LIB = os.path.join(PKG, "libtgr_embed.so")
PACK_LIB = os.path.join(PKG, "libtgr_pack.so")     # host-side dict tensorizer (plain C on the CPython API)
PACK_SRC = os.path.join(CSRC, "tgr_pack.c")
CC = os.environ.get("CC", "gcc")
SOURCES = ["tgr_util.cu", "tgr_fwd.cu", "tgr_mm.cu", "tgr_mm_tc.cu", "tgr_bwd.cu", "tgr_sort.cu", "tgr_reduce.cu", "tgr_route.cu", "tgr_symm.cu", "tgr_resident.cu", "tgr_factored.cu", "tgr_rows_ws.cu", "tgr_fact_step.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC",
         "-I", INCLUDE, "-I", CSRC, "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _stale(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    headers = [os.path.join(CSRC, h) for h in os.listdir(CSRC) if h.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "tgr_embed.h"))
    headers.append(os.path.abspath(__file__))
    bdir = os.path.join(CSRC, "build")
    os.makedirs(bdir, exist_ok=True)
    objs, jobs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(bdir, s[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
        return job, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for (src, obj), r in ex.map(compile_one, jobs):
                log = os.path.join(bdir, os.path.basename(src) + ".ptxas.log")
                with open(log, "w") as f:
                    f.write(r.stderr)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed on {src}")
                if verbose:
                    sys.stderr.write(r.stderr)
    if jobs or force or _stale(LIB, objs):
        r = subprocess.run([NVCC, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-lcudart"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    build_pack(force)
    return LIB


def build_pack(force: bool = False) -> str:
    """gcc -shared tgr_pack.c -> libtgr_pack.so (symbols of libpython resolve against the running interpreter)."""
    import sysconfig
    if force or _stale(PACK_LIB, [PACK_SRC, os.path.abspath(__file__)]):
        inc = sysconfig.get_paths()["include"]
        r = subprocess.run([CC, "-O2", "-Wall", "-shared", "-fPIC", "-I", inc, PACK_SRC, "-o", PACK_LIB],
                           capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("gcc failed on tgr_pack.c")
    return PACK_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
