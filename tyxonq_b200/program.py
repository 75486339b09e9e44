"""Device-side execution of compiled programs and the thin Python wrappers of the C ABI.

torch is used for device memory, streams and host<->device copies only; all 2^n-sized
arithmetic runs in libtyxonq_b200.so.  Every function raises ``TqbError`` when the library
or the GPU is missing -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .gates import LGate
from .planner import Program, TileConfig, compile_program, default_tile


def _dev_index(t: torch.Tensor) -> int:
    if not t.is_cuda:
        raise _lib.TqbError("state tensors must live on a CUDA device")
    return t.device.index if t.device.index is not None else torch.cuda.current_device()


def _prep(state: torch.Tensor) -> Tuple[int, int, int, int, int]:
    """-> (ptr, n, batch, dtype_code, stream) for a contiguous [2^n] or [batch, 2^n] complex tensor."""
    if not state.is_contiguous():
        raise _lib.TqbError("state tensor must be contiguous")
    dev = _dev_index(state)
    _lib.ensure_device(dev)
    dim = state.shape[-1]
    n = int(dim).bit_length() - 1
    if (1 << n) != dim:
        raise _lib.TqbError("last dimension must be a power of two")
    batch = 1 if state.dim() == 1 else int(state.shape[0])
    if state.dim() > 2:
        raise _lib.TqbError("state must be [2^n] or [batch, 2^n]")
    return state.data_ptr(), n, batch, _lib.dtype_code(state.dtype), _lib.current_stream_ptr(state.device)


def new_state(n: int, *, batch: int = 1, dtype: torch.dtype = torch.complex128, device: str | torch.device = "cuda",
              basis_index: int = 0, global_base: int = 0) -> torch.Tensor:
    """|basis_index> for every batch member (replaces init_statevector, statevector.py:19-25)."""
    device = torch.device(device)
    shape = (1 << n,) if batch == 1 else (batch, 1 << n)
    with torch.cuda.device(device):
        st = torch.empty(shape, dtype=dtype, device=device)
        ptr, n_, b_, dt, stream = _prep(st)
        _lib.check(_lib.load().tqb_init_basis(ptr, n_, b_, dt, global_base, basis_index, stream))
    return st


class DeviceProgram:
    """A compiled program resident on one device (gate descriptors + matrices uploaded once)."""

    def __init__(self, prog: Program, device: torch.device, dtype: torch.dtype) -> None:
        self.prog = prog
        self.device = torch.device(device)
        self.dtype = dtype
        self.h2d_bytes = 0
        self._passes = np.ascontiguousarray(prog.passes)
        self._gates_host = np.ascontiguousarray(prog.gates)   # host copy: lets the library pick specialised kernels
        with torch.cuda.device(self.device):
            g_host = torch.from_numpy(prog.gates.view(np.uint8).reshape(-1).copy()).pin_memory() if prog.gates.size else None
            self.gates_dev = (g_host.to(self.device, non_blocking=True) if g_host is not None
                              else torch.zeros(1, dtype=torch.uint8, device=self.device))
            self._g_host = g_host
            self.mats_host = torch.empty(max(1, prog.mats.size), dtype=dtype).pin_memory()
            self.mats_dev = torch.empty(max(1, prog.mats.size), dtype=dtype, device=self.device)
            self.h2d_bytes += prog.gates.nbytes
            self.upload_mats(prog.mats)

    def fill_host(self, mats: np.ndarray) -> None:
        """Write new matrices into the pinned staging buffer (host only; see ``upload``)."""
        src = torch.from_numpy(np.ascontiguousarray(mats, dtype=np.complex128))
        if src.numel():
            self.mats_host[: src.numel()].copy_(src)  # casts to complex64 when needed

    def upload(self) -> None:
        """Async H2D copy of the staging buffer on the current stream (capturable into a CUDA graph:
        a replay re-reads the pinned buffer, so new parameters only need ``fill_host``)."""
        with torch.cuda.device(self.device):
            self.mats_dev.copy_(self.mats_host, non_blocking=True)
        self.h2d_bytes += self.mats_host.numel() * self.mats_host.element_size()

    def upload_mats(self, mats: np.ndarray) -> None:
        """(Re)load the matrix buffer, e.g. for new variational parameters."""
        if np.asarray(mats).size:
            self.fill_host(mats)
            self.upload()

    def run(self, state: torch.Tensor, *, global_base: int = 0) -> torch.Tensor:
        if state.dtype != self.dtype:
            raise _lib.TqbError("state dtype does not match the program's dtype")
        with torch.cuda.device(self.device):
            ptr, n, batch, dt, stream = _prep(state)
            if n != self.prog.n:
                raise _lib.TqbError(f"program compiled for n={self.prog.n}, state has n={n}")
            if self.prog.n_passes:
                t = self.prog.tile
                _lib.check(_lib.load().tqb_run_passes2(
                    ptr, n, batch, dt, global_base, self._passes.ctypes.data, self.prog.n_passes,
                    self.gates_dev.data_ptr(), self._gates_host.ctypes.data, self.mats_dev.data_ptr(), t.threads,
                    t.ctas_per_sm, stream))
        return state


def apply_gates(state: torch.Tensor, gates: Sequence[LGate], *, tile: Optional[TileConfig] = None,
                global_base: int = 0) -> torch.Tensor:
    """Compile and run ``gates`` in place on ``state``."""
    if not gates:
        return state
    ptr, n, batch, dt, _ = _prep(state)
    tile = tile or default_tile(n, state.element_size(), batch)
    prog = compile_program(list(gates), n, tile, itemsize=state.element_size())
    DeviceProgram(prog, state.device, state.dtype).run(state, global_base=global_base)
    return state


# ---- reductions -------------------------------------------------------------------------
def norm2(state: torch.Tensor) -> torch.Tensor:
    ptr, n, batch, dt, stream = _prep(state)
    out = torch.empty(batch, dtype=torch.float64, device=state.device)
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_norm2(ptr, n, batch, dt, out.data_ptr(), stream))
    return out


def expect_z_bits(state: torch.Tensor) -> torch.Tensor:
    """[batch, n] with column p = <Z> on index bit p (qubit q is column n-1-q)."""
    ptr, n, batch, dt, stream = _prep(state)
    out = torch.empty((batch, max(n, 1)), dtype=torch.float64, device=state.device)
    if n == 0:
        return out[:, :0]
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_expect_z_bits(ptr, n, batch, dt, out.data_ptr(), stream))
    return out


def expect_zmasks(state: torch.Tensor, masks: Sequence[int] | torch.Tensor, *, global_base: int = 0) -> torch.Tensor:
    ptr, n, batch, dt, stream = _prep(state)
    if not isinstance(masks, torch.Tensor):
        masks = torch.from_numpy(np.asarray(masks, dtype=np.uint64).view(np.int64)).to(state.device)
    nm = int(masks.numel())
    out = torch.empty((batch, nm), dtype=torch.float64, device=state.device)
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_expect_zmasks(ptr, n, batch, dt, global_base, masks.data_ptr(), nm, out.data_ptr(), stream))
    return out


def inner(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """<a|b> per batch member as complex128."""
    ptr, n, batch, dt, stream = _prep(a)
    ptrb, nb, bb, dtb, _ = _prep(b)
    if (n, batch, dt) != (nb, bb, dtb):
        raise _lib.TqbError("inner: shape/dtype mismatch")
    out = torch.empty((batch, 2), dtype=torch.float64, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().tqb_inner(ptr, ptrb, n, batch, dt, out.data_ptr(), stream))
    return torch.view_as_complex(out)


def project_z(state: torch.Tensor, bit: int, keep: int) -> torch.Tensor:
    ptr, n, batch, dt, stream = _prep(state)
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_project_z(ptr, n, batch, dt, int(bit), int(keep), stream))
    return state


def scale(state: torch.Tensor, factor: float) -> torch.Tensor:
    ptr, n, batch, dt, stream = _prep(state)
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_scale(ptr, n, batch, dt, float(factor), stream))
    return state


def probabilities(state: torch.Tensor) -> torch.Tensor:
    ptr, n, batch, dt, stream = _prep(state)
    out = torch.empty(state.shape, dtype=torch.float64, device=state.device)
    with torch.cuda.device(state.device):
        _lib.check(_lib.load().tqb_probabilities(ptr, n, batch, dt, out.data_ptr(), stream))
    return out


def sample(state: torch.Tensor, uniforms: torch.Tensor) -> torch.Tensor:
    """Indices drawn with host-supplied uniforms: idx = #{i : cdf_i/cdf_last <= u} over the blocked CDF
    (replaces Generator.choice in engine.py:415).  uniforms: [shots] or [batch, shots] float64."""
    ptr, n, batch, dt, stream = _prep(state)
    u = uniforms.to(device=state.device, dtype=torch.float64).contiguous()
    shots = int(u.shape[-1])
    if u.numel() != batch * shots:
        raise _lib.TqbError("uniforms must be [shots] or [batch, shots]")
    nc = 1 << (n - min(n, 12))
    prefix = torch.empty((batch, nc + 1), dtype=torch.float64, device=state.device)
    # in-chunk partial sums every 128 amplitudes (n >= 12): a shot then scans <= 128 amplitudes of its chunk
    sub = torch.empty((batch, nc, 32), dtype=torch.float64, device=state.device) if n >= 12 else None
    idx = torch.empty(u.shape, dtype=torch.int64, device=state.device)
    with torch.cuda.device(state.device):
        lib = _lib.load()
        sp = sub.data_ptr() if sub is not None else None
        _lib.check(lib.tqb_cdf_chunks2(ptr, n, batch, dt, prefix.data_ptr(), sp, stream))
        _lib.check(lib.tqb_sample2(ptr, n, batch, dt, prefix.data_ptr(), sp, u.data_ptr(), shots, idx.data_ptr(), stream))
    return idx
