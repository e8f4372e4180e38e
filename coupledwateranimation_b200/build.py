"""Build libcwa_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Every .cu is compiled to its own object (in parallel, rebuilt only when it or a header changed) and the objects are linked
into the shared library; no relocatable device code is needed (kernels never call across translation units)."""
from __future__ import annotations

import os
import shutil
import subprocess
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libcwa_b200.so")
SOURCES = ["api.cu", "grid.cu", "wave.cu", "sph3.cu", "sph2.cu", "multi.cu", "slab.cu", "stencil1d.cu", "state.cu", "interop.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libcwa_b200.so")


def _gl_flags():
    """cuda_gl_interop.h includes <GL/gl.h>; machines without OpenGL headers (this image) get the two-typedef stand-in."""
    for d in ("/usr/include", "/usr/local/include", "/usr/include/x86_64-linux-gnu"):
        if os.path.exists(os.path.join(d, "GL", "gl.h")):
            return []
    return ["-I", os.path.join(CSRC, "gl_compat")]


def _headers():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [os.path.join(HERE, "..", "include", "cwa_b200.h")]


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cwa_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    hdr_t = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    for s in SOURCES:
        src, obj = os.path.join(CSRC, s), os.path.join(OBJ, s[:-3] + ".o")
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((s, [nvcc] + NVCC_FLAGS + _gl_flags() + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]))

    def run(job):
        return job[0], subprocess.run(job[1], capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 4))) as ex:
        for name, r in ex.map(run, jobs):
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {name}:\n" + r.stdout + r.stderr)
            if verbose:
                print(f"==== {name}\n{r.stderr}")
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC", "-o", LIB] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
