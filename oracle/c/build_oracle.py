"""Build the C restatement (oracle/c/sv_oracle.c) into oracle/c/libsv_oracle.so with gcc + OpenMP."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
SRC = HERE / "sv_oracle.c"
LIB = HERE / "libsv_oracle.so"


def build() -> Path:
    if LIB.exists() and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    subprocess.run(["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-std=c11", str(SRC), "-o", str(LIB), "-lm"],
                   check=True)
    return LIB


if __name__ == "__main__":
    print(build())
