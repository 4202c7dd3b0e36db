"""In-tree build of libgmlm_b200.so (plain nvcc, no torch headers: the library is a C ABI).

    python -m gmlm_b200.build [--force]

The shared object lands next to this file (``gmlm_b200/libgmlm_b200.so``) so that it
travels with the repo snapshot to the GPU box; it is git-ignored.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
OBJ = PKG / "build"
LIB = PKG / "libgmlm_b200.so"

NVCC_FLAGS = [
    "-std=c++20", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libgmlm_b200.so cannot be built (there is no CPU fallback)")
    return exe


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _deps_mtime() -> float:
    hdrs = list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    return max(p.stat().st_mtime for p in hdrs) if hdrs else 0.0


def _compile_one(src: Path, verbose: bool) -> Path:
    obj = OBJ / (src.stem + ".o")
    cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src.name}:\n{r.stdout}\n{r.stderr}")
    if verbose:
        (OBJ / (src.stem + ".ptxas.log")).write_text(r.stderr)
    return obj


def build(force: bool = False, verbose: bool = False) -> Path:
    OBJ.mkdir(exist_ok=True)
    srcs = _sources()
    dep_t = _deps_mtime()
    todo = []
    for s in srcs:
        o = OBJ / (s.stem + ".o")
        if force or not o.exists() or o.stat().st_mtime < max(s.stat().st_mtime, dep_t):
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            list(ex.map(lambda s: _compile_one(s, verbose), todo))
    objs = [OBJ / (s.stem + ".o") for s in srcs]
    if force or todo or not LIB.exists() or any(o.stat().st_mtime > LIB.stat().st_mtime for o in objs):
        cmd = [_nvcc(), "-shared", "-o", str(LIB), *map(str, objs), ]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    lib = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(lib)
