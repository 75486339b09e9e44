"""The C-ABI shared library loads without a GPU and exports every symbol include/*.h declares."""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]


def _declared_symbols():
    text = (ROOT / "include" / "tyxonq_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tqb_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_header_symbols():
    from tyxonq_b200 import _lib, build
    build.build_library()
    lib = _lib.load()
    syms = _declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tyxonq_b200.h but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert sorted(_lib.SIGNATURES) == syms
    assert lib.tqb_abi_version() == 2


def test_struct_layouts_match_header():
    from tyxonq_b200 import _lib

    class Gate(ctypes.Structure):
        _fields_ = [("kind", ctypes.c_int32), ("k", ctypes.c_int32), ("bits", ctypes.c_int8 * 8), ("sbits", ctypes.c_int8 * 8),
                    ("off_a", ctypes.c_uint32), ("off_b", ctypes.c_uint32), ("mat_off", ctypes.c_uint32),
                    ("mat_bstride", ctypes.c_uint32), ("zmask", ctypes.c_uint64)]

    class Pass(ctypes.Structure):
        _fields_ = [("m", ctypes.c_int32), ("L", ctypes.c_int32), ("gate_begin", ctypes.c_int32), ("n_gates", ctypes.c_int32),
                    ("max_dense_k", ctypes.c_int32), ("mat_begin", ctypes.c_int32), ("mat_count", ctypes.c_int32),
                    ("hb", ctypes.c_int8 * 16)]

    assert ctypes.sizeof(Gate) == _lib.GATE_DTYPE.itemsize == 48
    assert ctypes.sizeof(Pass) == _lib.PASS_DTYPE.itemsize == 44
    for name, _ in Gate._fields_:
        assert getattr(Gate, name).offset == _lib.GATE_DTYPE.fields[name][1], name
    for name, _ in Pass._fields_:
        assert getattr(Pass, name).offset == _lib.PASS_DTYPE.fields[name][1], name


def test_no_cpu_fallback_without_gpu():
    """Product entry points must fail loudly when no CUDA device is present."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tyxonq_b200 import TqbError, kernels
    with pytest.raises(TqbError):
        kernels.apply_1q_statevector(None, np.array([1, 0], dtype=complex), np.eye(2), 0, 1)


def test_product_never_imports_oracle():
    for py in (ROOT / "tyxonq_b200").rglob("*.py"):
        src = py.read_text()
        assert "import oracle" not in src and "from oracle" not in src, py
