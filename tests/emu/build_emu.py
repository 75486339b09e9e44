"""Build the host emulator of the tile kernel (test tooling; see tqb_emu.cpp)."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "libtqb_emu.so"
SRC = HERE / "tqb_emu.cpp"
DEPS = [SRC, HERE.parents[1] / "tyxonq_b200" / "csrc" / "tqb_core.cuh", HERE.parents[1] / "include" / "tyxonq_b200.h"]


def build() -> Path:
    if LIB.exists() and all(d.stat().st_mtime <= LIB.stat().st_mtime for d in DEPS):
        return LIB
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-x", "c++", str(SRC), "-o", str(LIB)], check=True)
    return LIB


if __name__ == "__main__":
    print(build())
