"""ctypes binding of libtyxonq_b200.so (the C ABI declared in include/tyxonq_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is not
available, every compute entry point raises.  Loading the library itself needs only
libcudart, so symbol checks work on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path
from typing import Optional

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libtyxonq_b200.so"
JIT_CACHE = PKG / "jit_cache"
if os.environ.get("TQB_LIB"):   # profiling builds only (tools/build_prof.sh): an alternative build of the same sources
    LIB_PATH = Path(os.environ["TQB_LIB"])

TQB_C64, TQB_C128, TQB_F64 = 0, 1, 2
GATE_DENSE, GATE_DIAG, GATE_PAIR, GATE_SWAP = 0, 1, 2, 3
GATE_MUX, GATE_CHAIN = 4, 5
MAX_GATE_BITS = 8
MAX_TILE_HIGH = 16
SCAN_BLOCK = 4096

# numpy mirrors of the ABI structs (layout checked against ctypes.sizeof below)
GATE_DTYPE = np.dtype([
    ("kind", "<i4"), ("k", "<i4"),
    ("bits", "i1", (MAX_GATE_BITS,)), ("sbits", "i1", (MAX_GATE_BITS,)),
    ("off_a", "<u4"), ("off_b", "<u4"), ("mat_off", "<u4"), ("mat_bstride", "<u4"),
    ("zmask", "<u8"),
], align=True)
PASS_DTYPE = np.dtype([
    ("m", "<i4"), ("L", "<i4"), ("gate_begin", "<i4"), ("n_gates", "<i4"), ("max_dense_k", "<i4"),
    ("mat_begin", "<i4"), ("mat_count", "<i4"),
    ("hb", "i1", (MAX_TILE_HIGH,)),
], align=True)
PAIR_STEP_DTYPE = np.dtype([
    ("k", "<i4"), ("slot", "<i4"), ("sbits", "i1", (MAX_GATE_BITS,)),
    ("off_a", "<u8"), ("off_b", "<u8"), ("zmask", "<u8"), ("c", "<f8"), ("s", "<f8"), ("scale", "<f8"),
], align=True)
VQE_OP_DTYPE = np.dtype([
    ("kind", "<i4"), ("param", "<i4"), ("xmask", "<u4"), ("zmask", "<u4"), ("bit0", "<i4"), ("bit1", "<i4"),
    ("mat_off", "<i4"), ("reserved", "<i4"), ("scale", "<f8"),
], align=True)
assert VQE_OP_DTYPE.itemsize == 40, VQE_OP_DTYPE.itemsize
assert PAIR_STEP_DTYPE.itemsize == 64, PAIR_STEP_DTYPE.itemsize
assert GATE_DTYPE.itemsize == 48, GATE_DTYPE.itemsize
assert PASS_DTYPE.itemsize == 44, PASS_DTYPE.itemsize

_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_u64 = C.c_uint64
_d = C.c_double

# name -> (restype, argtypes); must list every symbol include/tyxonq_b200.h declares
SIGNATURES = {
    "tqb_abi_version": (_i, []),
    "tqb_last_error": (C.c_char_p, []),
    "tqb_init": (_i, [_i]),
    "tqb_shutdown": (_i, [_i]),
    "tqb_launch_count": (_i64, []),
    "tqb_device_info": (_i, [_i, C.POINTER(_i), C.POINTER(_i)]),
    "tqb_init_basis": (_i, [_vp, _i, _i64, _i, _u64, _u64, _vp]),
    "tqb_run_passes": (_i, [_vp, _i, _i64, _i, _u64, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "tqb_run_passes2": (_i, [_vp, _i, _i64, _i, _u64, _vp, _i, _vp, _vp, _vp, _i, _i, _vp]),
    "tqb_workspace_slot": (_i, [_i, _i]),
    "tqb_set_jit": (_i, [_i]),
    "tqb_set_jit_cache": (_i, [C.c_char_p]),
    "tqb_jit_wait": (_i, []),
    "tqb_jit_shutdown": (_i, []),
    "tqb_jit_stats": (_i, [C.POINTER(_i64)]),
    "tqb_spec_source": (_i64, [_vp, _vp, _i, _i, _vp, _i64]),
    "tqb_spec_compile": (_i, [_vp, _vp, _i]),
    "tqb_set_tma": (_i, [_i]),
    "tqb_norm2": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tqb_expect_z_bits": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tqb_expect_zmasks": (_i, [_vp, _i, _i64, _i, _u64, _vp, _i, _vp, _vp]),
    "tqb_expect_pauli_sum": (_i, [_vp, _i, _i64, _i, _u64, _vp, _vp, _i, _vp, _vp, _vp, _vp]),
    "tqb_expect_pauli_tiled": (_i, [_vp, _i, _i64, _i, _u64, _vp, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "tqb_apply_pauli_sum": (_i, [_vp, _vp, _i, _i64, _i, _u64, _vp, _vp, _i, _vp, _vp, _vp]),
    "tqb_inner": (_i, [_vp, _vp, _i, _i64, _i, _vp, _vp]),
    "tqb_transition_1q": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp]),
    "tqb_grad_pair": (_i, [_vp, _vp, _i, _i, _vp, _d, _vp, _i, _vp]),
    "tqb_grad_dense": (_i, [_vp, _vp, _i, _i, _i, _vp, _vp, _d, _vp, _i, _vp]),
    "tqb_pair_sweep": (_i, [_vp, _vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp]),
    "tqb_vqe_resident": (_i, [_i, _vp, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i64, _vp, _vp]),
    "tqb_project_z": (_i, [_vp, _i, _i64, _i, _i, _i, _vp]),
    "tqb_scale": (_i, [_vp, _i, _i64, _i, _d, _vp]),
    "tqb_probabilities": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tqb_cdf_chunks": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tqb_sample": (_i, [_vp, _i, _i64, _i, _vp, _vp, _i64, _vp, _vp]),
    "tqb_cdf_chunks2": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp]),
    "tqb_sample2": (_i, [_vp, _i, _i64, _i, _vp, _vp, _vp, _i64, _vp, _vp]),
    "tqb_reduced_1q": (_i, [_vp, _i, _i64, _i, _i, _vp, _vp]),
    "tqb_dm_diag": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tqb_expval_from_samples": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "tqb_chunk_totals": (_i, [_vp, _i, _i64, _i, _vp, _vp]),
    "tqb_chunk_prefix": (_i, [_vp, _i64, _i64, _vp, _vp]),
    "tqb_sample_shard": (_i, [_vp, _i, _i, _vp, _i64, _i64, _i64, _vp, _i64, _vp, _vp]),
    "tqb_copy": (_i, [_vp, _vp, _i64, _i, _vp]),
}


class PauliLayout(C.Structure):
    """tqb_pauli_layout (include/tyxonq_b200.h)."""
    _fields_ = [("m", C.c_int32), ("L", C.c_int32), ("group_begin", C.c_int32), ("n_groups", C.c_int32), ("n_terms", C.c_int32),
                ("hb", C.c_int8 * MAX_TILE_HIGH)]


assert C.sizeof(PauliLayout) == 36


class TqbError(RuntimeError):
    pass


_lib: Optional[C.CDLL] = None
_inited_devices: set = set()


def load() -> C.CDLL:
    """Load the shared library (no GPU needed) and bind every ABI symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise TqbError(
            f"{LIB_PATH} is missing: build it with `python -m tyxonq_b200.build` "
            "(tyxonq_b200 has no CPU fallback)")
    _find_nvrtc()
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drift apart
        fn.restype = res
        fn.argtypes = args
    if lib.tqb_abi_version() != 2:
        raise TqbError("libtyxonq_b200.so ABI version mismatch")
    # specialised pass kernels (csrc/tqb_jit.cu): cubin cache next to the library; TQB_JIT = 0 off / 1 async / 2 sync
    lib.tqb_set_jit_cache(os.environ.get("TQB_JIT_CACHE", str(JIT_CACHE)).encode())
    if os.environ.get("TQB_JIT"):
        lib.tqb_set_jit(int(os.environ["TQB_JIT"]))
    if os.environ.get("TQB_TENSOR_TMA"):   # 0 = stage tiles with one bulk copy per run instead of one tensor copy per tile
        lib.tqb_set_jit(512 + int(os.environ["TQB_TENSOR_TMA"]))
    if os.environ.get("TQB_SPEC_LOOP"):    # code shape of the specialised kernels: 0 unrolled, 1 looped, 2 by dtype (default)
        lib.tqb_set_jit(1024 + int(os.environ["TQB_SPEC_LOOP"]))
    import atexit
    atexit.register(lib.tqb_jit_shutdown)   # before interpreter teardown: no compilation thread outlives the process
    _lib = lib
    return lib


def _find_nvrtc() -> None:
    """Point the library's dlopen at an NVRTC (the CUDA toolkit's, else the wheel torch depends on)."""
    if os.environ.get("TQB_NVRTC"):
        return
    cands = ["/usr/local/cuda/lib64/libnvrtc.so.12"]
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.cuda_nvrtc")
        if spec and spec.submodule_search_locations:
            cands.append(str(Path(list(spec.submodule_search_locations)[0]) / "lib" / "libnvrtc.so.12"))
    except Exception:
        pass
    for c in cands:
        if os.path.exists(c):
            os.environ["TQB_NVRTC"] = c
            return


def jit_stats() -> dict:
    out = (_i64 * 4)()
    load().tqb_jit_stats(out)
    return {"spec_launches": int(out[0]), "compiles": int(out[1]), "disk_hits": int(out[2]), "shapes": int(out[3])}


def check(rc: int) -> None:
    if rc != 0:
        msg = load().tqb_last_error()
        raise TqbError(msg.decode() if msg else f"tqb error {rc}")


def ensure_device(device_index: int) -> None:
    """tqb_init for a CUDA device (allocates the reduction workspace once)."""
    if device_index in _inited_devices:
        return
    import torch
    if not torch.cuda.is_available():
        raise TqbError("tyxonq_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    check(load().tqb_init(int(device_index)))
    if os.environ.get("TQB_TMA"):  # tuning knob: 0 = LDG/STG, 2 / 3 = TMA ring depth, 1 = auto
        load().tqb_set_tma(int(os.environ["TQB_TMA"]))
    if os.environ.get("TQB_LEAN"):  # tuning knob: 0 = never use the lean kernel variant
        load().tqb_set_tma(512 + int(os.environ["TQB_LEAN"]))
    if os.environ.get("TQB_DBG"):  # profiling only: 1 = no gate arithmetic, 2 = no bulk loads, 4 = no bulk stores
        load().tqb_set_tma(256 + int(os.environ["TQB_DBG"]))
        load().tqb_set_jit(256 + int(os.environ["TQB_DBG"]))
    _inited_devices.add(device_index)


def launch_count() -> int:
    return int(load().tqb_launch_count())


def dtype_code(torch_dtype) -> int:
    import torch
    if torch_dtype == torch.complex128:
        return TQB_C128
    if torch_dtype == torch.complex64:
        return TQB_C64
    raise TqbError(f"unsupported state dtype {torch_dtype}")


def current_stream_ptr(device) -> int:
    import torch
    return int(torch.cuda.current_stream(device).cuda_stream)
