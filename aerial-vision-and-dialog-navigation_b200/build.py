"""In-tree build of libavdn.so (sm_100a only) with nvcc.

``python -m avdn_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU.  The shared object is written next to the sources so that it
travels with the tree; nothing is JIT-cached elsewhere.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(CSRC, "libavdn.so")
INCLUDE = os.path.abspath(os.path.join(HERE, "..", "include"))

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-I", INCLUDE,
          "--expt-relaxed-constexpr", "-Xptxas", "-v"]

# per-file extra flags; render.cu must not contract a*b+c into an FMA because its
# float64 arithmetic has to round exactly like OpenCV's scalar code.
SOURCES = {
    "common.cu": [],
    "render.cu": ["-fmad=false"],
    "gemm.cu": [],
    "trunk.cu": [],
    "conv0.cu": [],
    "conv0_tc.cu": [],
    "conv3_halo.cu": [],
    "lstm.cu": ["-fmad=false"],
    "encoder.cu": [],
    "agent.cu": [],
    "bert.cu": [],
}


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; libavdn.so cannot be built")


def _digest(paths, flags) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(INCLUDE, "avdn.h"))
    objs, relink = [], force or not os.path.exists(LIB)
    logs = []
    jobs = []
    for src, extra in SOURCES.items():
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        stamp = obj + ".sha"
        flags = ARCH + COMMON + extra
        dig = _digest([path] + headers, flags)
        old = open(stamp).read() if os.path.exists(stamp) else ""
        if force or old != dig or not os.path.exists(obj):
            jobs.append((src, [nvcc] + flags + ["-c", path, "-o", obj], stamp, dig))
        objs.append(obj)
    if jobs:
        # the translation units are independent: compile them side by side (gemm.cu alone takes ~1 minute)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1)) as ex:
            results = list(ex.map(lambda j: subprocess.run(j[1], capture_output=True, text=True), jobs))
        for (src, cmd, stamp, dig), r in zip(jobs, results):
            logs.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
            if r.returncode != 0:
                sys.stderr.write(logs[-1])
                raise RuntimeError(f"nvcc failed on {src}")
            with open(stamp, "w") as f:
                f.write(dig)
        relink = True
    if relink:
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs
        r = subprocess.run(cmd, capture_output=True, text=True)
        logs.append(f"$ {' '.join(cmd)}\n{r.stdout}{r.stderr}")
        if r.returncode != 0:
            sys.stderr.write(logs[-1])
            raise RuntimeError("link of libavdn.so failed")
    with open(os.path.join(OBJ, "build.log"), "a" if not force else "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
