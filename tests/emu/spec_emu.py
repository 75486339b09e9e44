"""Host emulation of the specialised pass kernels (test tooling only; see spec_emu.cpp).

For every pass of a Program the library's own generator (tqb_spec_source: no GPU needed) prints the pass's
compile-time constants; this module compiles spec_emu.cpp against them with g++ (one shared object per distinct pass
shape, cached under tests/emu/_spec/) and runs the pass on a host array.  Passes that are not eligible for
specialisation run through the generic emulator (emu.py), exactly like tqb_run_passes2 falls back on the GPU.
"""
from __future__ import annotations

import ctypes as C
import hashlib
import subprocess
from pathlib import Path

import numpy as np

from tyxonq_b200 import _lib

HERE = Path(__file__).resolve().parent
CACHE = HERE / "_spec"
SRC = HERE / "spec_emu.cpp"
TEMPLATE = HERE.parents[1] / "tyxonq_b200" / "csrc" / "tqb_spec.cuh"


def spec_header(passes: np.ndarray, pi: int, gates: np.ndarray, dtype_code: int):
    """Generated constants of pass ``pi`` (None when the pass is not eligible)."""
    lib = _lib.load()
    buf = C.create_string_buffer(1 << 16)
    one = np.ascontiguousarray(passes[pi:pi + 1])
    r = lib.tqb_spec_source(one.ctypes.data, gates.ctypes.data, dtype_code, 0, buf, 1 << 16)
    return None if r < 0 else buf.value.decode()


def _build(header: str) -> C.CDLL:
    CACHE.mkdir(exist_ok=True)
    stamp = f"{SRC.stat().st_mtime_ns}:{TEMPLATE.stat().st_mtime_ns}"
    key = hashlib.sha1((header + stamp).encode()).hexdigest()[:20]
    so = CACHE / f"spec_{key}.so"
    if not so.exists():
        hdr = CACHE / f"spec_{key}.h"
        hdr.write_text(header)
        subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-include", str(hdr), str(SRC), "-o", str(so)],
                       check=True)
    lib = C.CDLL(str(so))
    lib.spec_emu_run.restype = C.c_int
    lib.spec_emu_run.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_ulonglong, C.c_void_p, C.c_void_p, C.c_void_p]
    return lib


def _ext_positions(header: str, pass_row, gates: np.ndarray) -> np.ndarray:
    """Outside-the-tile bit positions in slot order: first appearance over the pass's gates (the generator's rule)."""
    ext = []
    g0, ng = int(pass_row["gate_begin"]), int(pass_row["n_gates"])
    for g in gates[g0:g0 + ng]:
        kind, k = int(g["kind"]), int(g["k"])
        if kind == _lib.GATE_MUX:
            codes = [g["bits"][1]]
        elif kind == _lib.GATE_CHAIN:
            codes = [g["bits"][k]] + ([g["bits"][k + 1 + j] for j in range(int(g["off_b"]) & 3)] if int(g["off_a"]) >= 4 else [])
        elif kind == _lib.GATE_DIAG:
            codes = list(g["bits"][:k])
        else:
            codes = []
        for c in codes:
            u = int(c) & 0xff
            if 64 <= u < 127 and (u - 64) not in ext:
                ext.append(u - 64)
    out = np.zeros(8, dtype=np.int8)
    out[:len(ext)] = ext
    return out


def run_program_spec_emulated(prog, state: np.ndarray, *, batch: int = 1, global_base: int = 0):
    """-> (state after the program, number of passes that ran through a specialised-kernel emulation)."""
    from .emu import run_program_emulated
    st = np.ascontiguousarray(state).copy()
    code = 1 if st.dtype == np.complex128 else 0
    mats = np.ascontiguousarray(prog.mats.astype(st.dtype))
    passes = np.ascontiguousarray(prog.passes)
    gates = np.ascontiguousarray(prog.gates)
    n_spec = 0
    for pi in range(len(passes)):
        hdr = spec_header(passes, pi, gates, code) if prog.n > int(passes[pi]["m"]) else None
        if hdr is None:
            import copy
            sub = copy.copy(prog)
            sub.passes = passes[pi:pi + 1]
            st = run_program_emulated(sub, st, batch=batch, global_base=global_base, threads=128)
            continue
        lib = _build(hdr)
        ps = passes[pi]
        hb = np.ascontiguousarray(ps["hb"])
        ext = _ext_positions(hdr, ps, gates)
        m0 = mats[int(ps["mat_begin"]):]
        m0 = np.ascontiguousarray(m0)
        rc = lib.spec_emu_run(st.ctypes.data, prog.n, batch, global_base, hb.ctypes.data, ext.ctypes.data, m0.ctypes.data)
        assert rc == 0
        n_spec += 1
    return st, n_spec
