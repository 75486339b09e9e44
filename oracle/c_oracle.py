"""ctypes front-end of the C restatement (oracle/c/sv_oracle.c)  --  TEST INFRASTRUCTURE ONLY.

Interprets op tuples like ``sv_oracle.evolve_ops`` (same gate matrices, same skipped-op quirks)
but applies the gates with the OpenMP C loops, in place, so 26-30 qubit states can be timed on the
host cores.  Used by tests/test_oracle_c.py and by bench.py's cpu_baseline / --impl reference legs.
"""
from __future__ import annotations

import ctypes as C
from typing import Any, Iterable, Sequence

import numpy as np

from . import sv_oracle as O
from .c.build_oracle import build

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(str(build()))
        for suf in ("c128", "c64"):
            getattr(L, f"orc_apply_1q_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
            getattr(L, f"orc_apply_2q_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
            getattr(L, f"orc_expect_z_{suf}").argtypes = [C.c_void_p, C.c_int, C.c_int]
            getattr(L, f"orc_expect_z_{suf}").restype = C.c_double
            getattr(L, f"orc_init_zero_{suf}").argtypes = [C.c_void_p, C.c_int]
        L.orc_max_threads.restype = C.c_int
        L.orc_set_threads.restype = C.c_int
        L.orc_set_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def max_threads() -> int:
    return int(lib().orc_max_threads())


def set_threads(n: int) -> int:
    return int(lib().orc_set_threads(int(n)))


def _suf(psi: np.ndarray) -> str:
    return "c128" if psi.dtype == np.complex128 else "c64"


def new_state(n: int, dtype=np.complex128) -> np.ndarray:
    psi = np.empty(1 << n, dtype=dtype)
    getattr(lib(), f"orc_init_zero_{_suf(psi)}")(psi.ctypes.data, n)
    return psi


def apply_1q(psi: np.ndarray, g: np.ndarray, q: int, n: int) -> None:
    g = np.ascontiguousarray(g, dtype=np.complex128)
    getattr(lib(), f"orc_apply_1q_{_suf(psi)}")(psi.ctypes.data, n, n - 1 - q, g.ctypes.data)


def apply_2q(psi: np.ndarray, g: np.ndarray, q0: int, q1: int, n: int) -> None:
    if q0 == q1:
        return
    g = np.ascontiguousarray(g, dtype=np.complex128)
    getattr(lib(), f"orc_apply_2q_{_suf(psi)}")(psi.ctypes.data, n, n - 1 - q0, n - 1 - q1, g.ctypes.data)


def expect_z(psi: np.ndarray, q: int, n: int) -> float:
    return float(getattr(lib(), f"orc_expect_z_{_suf(psi)}")(psi.ctypes.data, n, n - 1 - q))


def apply_ops(psi: np.ndarray, n: int, ops: Iterable[Sequence[Any]], mode: str = "run") -> int:
    """Apply gate ops in place; returns the number of gates applied."""
    cnt = 0
    for op in ops:
        nm = op[0]
        if nm in O._ONE_Q_FIXED:
            apply_1q(psi, O._ONE_Q_FIXED[nm](), int(op[1]), n)
        elif nm in O._ONE_Q_PARAM:
            apply_1q(psi, O._ONE_Q_PARAM[nm](float(op[2])), int(op[1]), n)
        elif nm in O._TWO_Q_FIXED:
            apply_2q(psi, O._TWO_Q_FIXED[nm](), int(op[1]), int(op[2]), n)
        elif nm in O._TWO_Q_PARAM:
            apply_2q(psi, O._TWO_Q_PARAM[nm](float(op[3])), int(op[1]), int(op[2]), n)
        elif nm == "cry" and mode == "run":
            apply_2q(psi, O.gate_cry_4x4(float(op[3])), int(op[1]), int(op[2]), n)
        else:
            continue
        cnt += 1
    return cnt
