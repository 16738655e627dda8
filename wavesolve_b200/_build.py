"""Build the CUDA library in-tree with nvcc for sm_100a (no JIT cache: the built
``libwavesolve_b200.so`` travels with the repository snapshot to the GPU box)."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libwavesolve_b200.so")
BUILD = os.path.join(PKG, "csrc", "_build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3"]
# translation unit -> extra flags.  The assembly kernels must reproduce numpy's separate
# multiply / add roundings, so FMA contraction is disabled for that file only.
SOURCES = {
    "ws_pattern.cu": [],
    "ws_assemble.cu": ["-fmad=false"],
    "ws_spmv.cu": [],
    "ws_symbolic.cu": [],
    "ws_symbolic_dev.cu": [],
    "ws_factor.cu": [],
    "ws_eig.cu": [],
    "ws_interp.cu": ["-fmad=false"],
    "ws_overlap.cu": [],
    "ws_hostcopy.cu": [],
}


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every CUDA source for sm_100a and link ``libwavesolve_b200.so``."""
    nvcc = _nvcc()
    os.makedirs(BUILD, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(PKG), "include", "wavesolve_b200.h"))
    headers.append(os.path.abspath(__file__))
    objs, procs = [], []
    for src, extra in SOURCES.items():
        s = os.path.join(CSRC, src)
        o = os.path.join(BUILD, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _stale(o, [s] + headers):
            cmd = [nvcc] + ARCH + COMMON + extra + ["-c", s, "-o", o]
            if verbose:
                print(" ".join(cmd), flush=True)
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    failed = []
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            failed.append(src + ":\n" + out.decode(errors="replace"))
        elif verbose and out:
            print(out.decode(errors="replace"))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(failed))
    if force or procs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        if verbose:
            print(" ".join(cmd), flush=True)
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout.decode(errors="replace"))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
