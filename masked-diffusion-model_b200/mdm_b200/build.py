"""Build libmdm_sm100.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python masked-diffusion-model_b200/mdm_b200/build.py [--force] [--verbose]

Each csrc/*.cu is compiled to build/<name>.o (in parallel, rebuilt only when the source or a
header is newer), then linked with a static cudart so the library has no libcuda link-time
dependency (the driver entry point for tensor-map encoding is resolved at run time)."""
from __future__ import annotations

import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmdm_sm100.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
FLAGS += [f"-D{d}" for d in os.environ.get("MDM_NVCC_DEFINES", "").split() if d]   # e.g. MDM_IGEMM_DEBUG_WAIT (debug builds only)


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers_mtime():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(HERE, "..", "..", "include", "mdm.h"))
    return max(os.path.getmtime(h) for h in hs)


def _compile(src, force, verbose):
    obj = os.path.join(BUILD, src[:-3] + ".o")
    path = os.path.join(CSRC, src)
    if (not force and os.path.exists(obj)
            and os.path.getmtime(obj) > max(os.path.getmtime(path), _headers_mtime())):
        return obj, ""
    cmd = [NVCC, *FLAGS, "-c", path, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}\n{p.stderr}")
    log = p.stderr if verbose else ""
    with open(obj + ".ptxas.txt", "w") as f:
        f.write(p.stderr)
    return obj, log


def build(force=False, verbose=False):
    os.makedirs(BUILD, exist_ok=True)
    srcs = _sources()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(lambda s: _compile(s, force, verbose), srcs))
    objs = [o for o, _ in res]
    for _, log in res:
        if log:
            print(log)
    if (force or not os.path.exists(LIB)
            or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-cudart", "static",
               "-gencode", "arch=compute_100a,code=sm_100a"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            raise RuntimeError(f"link failed:\n{p.stdout}\n{p.stderr}")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
