"""Build libtyxonq_b200.so (sm_100a) in-tree with nvcc.

The shared object is git-ignored but travels to the GPU box with the snapshot, so the
GPU side never compiles.  ``python -m tyxonq_b200.build`` or ``__graft_entry__.build()``.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libtyxonq_b200.so"
SOURCES = ["tqb_tile.cu", "tqb_reduce.cu", "tqb_jit.cu", "tqb_vqe.cu"]
HEADERS = ["tqb_core.cuh", "tqb_host.h", "tqb_spec.cuh", "../../include/tyxonq_b200.h"]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "-Xptxas=-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libtyxonq_b200.so")


def embed_spec_template() -> Path:
    """csrc/tqb_spec.cuh -> csrc/tqb_spec_src.inc (a C++ raw string literal): the kernel text tqb_jit.cu hands to NVRTC."""
    src = (CSRC / "tqb_spec.cuh").read_text()
    assert ')TQBSPEC"' not in src
    out = CSRC / "tqb_spec_src.inc"
    # (string literals are limited to 64 KiB by some front ends: split into adjacent literals)
    parts = [src[i:i + 12000] for i in range(0, len(src), 12000)]
    text = "\n".join('R"TQBSPEC(' + p + ')TQBSPEC"' for p in parts) + "\n"
    if not out.exists() or out.read_text() != text:
        out.write_text(text)
    return out


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = [CSRC / s for s in SOURCES] + [(CSRC / h).resolve() for h in HEADERS]
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    nvcc = _nvcc()
    embed_spec_template()
    objs = []
    build_dir = PKG / "build"
    build_dir.mkdir(exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = build_dir / (src + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        (build_dir / (src + ".ptxas.log")).write_text(out)
    cmd = [nvcc, "-shared", "-o", str(LIB), *map(str, objs), "-lcudart", "-ldl"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link of libtyxonq_b200.so failed")
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv))
