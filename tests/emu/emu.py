"""ctypes wrapper of the host emulator (test tooling only)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .build_emu import build


def run_program_emulated(prog, state: np.ndarray, *, batch: int = 1, global_base: int = 0, threads: int = 64) -> np.ndarray:
    """Run a planner.Program on a host array through the emulated tile kernel (in place on a copy)."""
    lib = C.CDLL(str(build()))
    lib.tqb_emu_run_passes.restype = C.c_int
    lib.tqb_emu_run_passes.argtypes = [C.c_void_p, C.c_int, C.c_longlong, C.c_int, C.c_ulonglong, C.c_void_p, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_int]
    st = np.ascontiguousarray(state).copy()
    dtype = 1 if st.dtype == np.complex128 else 0
    mats = np.ascontiguousarray(prog.mats.astype(st.dtype))
    passes = np.ascontiguousarray(prog.passes)
    gates = np.ascontiguousarray(prog.gates)
    rc = lib.tqb_emu_run_passes(st.ctypes.data, prog.n, batch, dtype, global_base, passes.ctypes.data, len(passes),
                                gates.ctypes.data, mats.ctypes.data, threads)
    if rc:
        raise RuntimeError(f"emulator rc={rc}")
    return st
