"""In-tree build of the C-ABI library ``bitdance_b200/_C/libbitdance_b200.so`` for sm_100a.

nvcc cross-compiles without a GPU; the resulting .so travels to the GPU box with the repo snapshot. Objects are
rebuilt only when a source or header is newer (mtime), so repeated ``build()`` calls are cheap.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_C")
LIB_PATH = os.path.join(OUT_DIR, "libbitdance_b200.so")
INCLUDE = os.path.join(os.path.dirname(HERE), "include")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "--expt-relaxed-constexpr",
    "-I", INCLUDE,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found; cannot build the sm_100a extension")


def _sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(INCLUDE, f) for f in os.listdir(INCLUDE)]
    return hs


def _compile_one(nvcc: str, src: str, obj: str, log: str) -> tuple[str, int, str]:
    cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
    p = subprocess.run(cmd, capture_output=True, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout + p.stderr)
    return src, p.returncode, p.stdout + p.stderr


def build(verbose: bool = False, force: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OUT_DIR, exist_ok=True)
    hdr_mtime = max(os.path.getmtime(h) for h in _headers())
    jobs = []
    objs = []
    for src in _sources():
        base = os.path.splitext(os.path.basename(src))[0]
        obj = os.path.join(OUT_DIR, base + ".o")
        objs.append(obj)
        stale = force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_mtime)
        if stale:
            jobs.append((src, obj, os.path.join(OUT_DIR, base + ".ptxas.log")))
    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            futs = [ex.submit(_compile_one, nvcc, *j) for j in jobs]
            for fu in futs:
                src, rc, out = fu.result()
                if verbose or rc != 0:
                    sys.stderr.write(out)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed for {src}")
    if jobs or not os.path.exists(LIB_PATH):
        cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        p = subprocess.run(cmd, capture_output=True, text=True)
        if p.returncode != 0:
            sys.stderr.write(p.stdout + p.stderr)
            raise RuntimeError("link failed")
    return LIB_PATH


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
