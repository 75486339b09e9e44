"""Lazy host views of device-resident results.

The reference's simulator driver ends a ``shots == 0`` run by simulating the circuit a second time and copying the
whole state to the host twice -- ``psi = eng.state(c); prob = np.abs(np.asarray(psi)) ** 2; statevec = np.asarray(psi)``
(devices/simulators/driver.py:115-126).  At 30 qubits that is 16 GiB over PCIe plus a 8 GiB probability vector nobody
asked for.  ``LazyHostArray`` stands in for those two entries of the result dict: it knows its shape and dtype, indexes
and slices by copying only what is asked for, and becomes a real numpy array the moment numpy wants one
(``np.asarray(x)``, ufuncs, ``x.sum()`` ...), so code written against the reference's result dict keeps working.
"""
from __future__ import annotations

from typing import Any, Callable, Optional, Tuple

import numpy as np
import torch


class LazyHostArray:
    """A 1-D array that lives on the device until the host really needs it."""

    __array_priority__ = 100.0

    def __init__(self, device_tensor: torch.Tensor, *, transform: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 dtype: Any = None) -> None:
        self._t = device_tensor
        self._f = transform
        self._host: Optional[np.ndarray] = None
        self.shape: Tuple[int, ...] = tuple(device_tensor.shape)
        self.dtype = np.dtype(dtype if dtype is not None else (np.complex128 if device_tensor.is_complex() else np.float64))
        self.ndim = len(self.shape)
        self.size = int(np.prod(self.shape)) if self.shape else 1

    @property
    def device_tensor(self) -> torch.Tensor:
        """The underlying CUDA tensor (before the transform): stay on the device with this."""
        return self._t

    @property
    def materialized(self) -> bool:
        return self._host is not None

    def _piece(self, t: torch.Tensor) -> np.ndarray:
        if self._f is not None:
            t = self._f(t)
        return t.detach().cpu().numpy().astype(self.dtype, copy=False)

    def __array__(self, dtype: Any = None, copy: Any = None) -> np.ndarray:
        if self._host is None:
            self._host = self._piece(self._t)
        return self._host if dtype is None else self._host.astype(dtype, copy=False)

    def __len__(self) -> int:
        return self.shape[0] if self.shape else 0

    def __getitem__(self, idx: Any) -> Any:
        if self._host is not None:
            return self._host[idx]
        if isinstance(idx, (int, np.integer, slice)):
            out = self._piece(self._t[idx])   # only the requested elements cross PCIe
            return out[()] if out.ndim == 0 else out
        return np.asarray(self)[idx]

    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):   # noqa: ANN001
        args = [np.asarray(x) if isinstance(x, LazyHostArray) else x for x in inputs]
        return getattr(ufunc, method)(*args, **kwargs)

    def __getattr__(self, name: str) -> Any:   # sum(), reshape(), real ... : numpy semantics after materialisation
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(np.asarray(self), name)

    def __repr__(self) -> str:
        state = "host" if self._host is not None else f"device {self._t.device}"
        return f"LazyHostArray(shape={self.shape}, dtype={self.dtype}, {state})"
