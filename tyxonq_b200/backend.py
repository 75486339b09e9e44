"""Seam B3 of SURVEY.md section 8b: an ``ArrayBackend`` for ``tq.set_backend(instance)``.

The reference keeps one process-global numerics backend (numerics/__init__.py:20-36, context.py:14-15) and
every kernel does ``K = backend or get_backend(None)`` (numerics/api.py:19-121 is the protocol).  ``B200Backend``
is that object for this package: arrays are torch tensors on the CUDA device, and ``value_and_grad`` differentiates
through ``StatevectorEngine.state`` with the adjoint sweep of ``autograd.py`` (the reference's torch backend keeps
one 2^n tensor per gate on the tape, pytorch_backend.py:446-564).  torch is plumbing here: the statevector work
itself runs in libtyxonq_b200.so once ``install()`` has rebound the engine and the kernel functions.

    import tyxonq as tq, tyxonq_b200
    tyxonq_b200.install()                                # seams B1 + B2
    tq.set_backend(tyxonq_b200.B200Backend())            # seam B3: Circuit.state() now returns device tensors

There is no CPU mode: constructing the backend without a CUDA device raises.
"""
from __future__ import annotations

import warnings
from typing import Any, Callable, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib


class B200Backend:
    name = "b200"
    available = True

    complex64 = torch.complex64
    complex128 = torch.complex128
    float32 = torch.float32
    float64 = torch.float64
    int8 = torch.int8
    int32 = torch.int32
    int64 = torch.int64
    bool = torch.bool
    int = torch.int64

    dtypestr = "complex128"
    rdtypestr = "float64"

    def __init__(self, device: str | torch.device | None = None) -> None:
        if not torch.cuda.is_available():
            raise _lib.TqbError("B200Backend needs a CUDA device (sm_100a); tyxonq_b200 has no CPU fallback")
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        _lib.ensure_device(self.device.index or 0)

    # ---- dtype handling (pytorch_backend.py:35-75) ----------------------------------------
    def set_dtype(self, dtype_str: str) -> Tuple[Any, Any]:
        if dtype_str == "complex64":
            self.dtypestr, self.rdtypestr = "complex64", "float32"
            return self.complex64, self.float32
        if dtype_str == "complex128":
            self.dtypestr, self.rdtypestr = "complex128", "float64"
            return self.complex128, self.float64
        raise ValueError(f"Unsupported dtype: {dtype_str}. Use 'complex64' or 'complex128'.")

    @staticmethod
    def _dt(dtype: Any) -> Optional[torch.dtype]:
        if dtype is None or isinstance(dtype, torch.dtype):
            return dtype
        if isinstance(dtype, str):
            return getattr(torch, dtype, None)
        table = {np.float32: torch.float32, np.float64: torch.float64, np.complex64: torch.complex64,
                 np.complex128: torch.complex128, np.int32: torch.int32, np.int64: torch.int64, np.int8: torch.int8,
                 np.bool_: torch.bool, float: torch.float64, complex: torch.complex128, int: torch.int64, bool: torch.bool}
        try:
            return table.get(dtype) or table.get(np.dtype(dtype).type)
        except TypeError:
            return None

    # ---- creation / conversion: requires_grad must survive (pytorch_backend.py:77-128) ----
    def array(self, data: Any, dtype: Any | None = None) -> Any:
        td = self._dt(dtype)
        if torch.is_tensor(data):
            out = data.to(device=self.device, dtype=td) if td is not None else data.to(self.device)
            if data.requires_grad and not out.requires_grad:
                out = out.requires_grad_(True)
            return out
        return torch.as_tensor(data, dtype=td, device=self.device)

    def asarray(self, data: Any, dtype: Any | None = None) -> Any:
        return self.array(data, dtype)

    def to_numpy(self, data: Any) -> np.ndarray:
        return data.detach().cpu().numpy() if torch.is_tensor(data) else np.asarray(data)

    # ---- algebra ----------------------------------------------------------------------------
    def matmul(self, a: Any, b: Any) -> Any:
        return self.asarray(a) @ self.asarray(b)

    def dot(self, a: Any, b: Any) -> Any:
        a, b = self.asarray(a), self.asarray(b)
        if a.dtype != b.dtype:
            dt = torch.promote_types(a.dtype, b.dtype)
            a, b = a.to(dt), b.to(dt)
        return a @ b

    def einsum(self, subscripts: str, *operands: Any) -> Any:
        return torch.einsum(subscripts, *[self.asarray(o) for o in operands])

    def reshape(self, a: Any, shape: Any) -> Any:
        return torch.reshape(self.asarray(a), tuple(shape) if not isinstance(shape, int) else (shape,))

    def moveaxis(self, a: Any, source: int, destination: int) -> Any:
        return torch.movedim(a, source, destination)

    def sum(self, a: Any, axis: int | None = None) -> Any:
        return torch.sum(a) if axis is None else torch.sum(a, dim=axis)

    def mean(self, a: Any, axis: int | None = None) -> Any:
        return torch.mean(a) if axis is None else torch.mean(a, dim=axis)

    def abs(self, a: Any) -> Any:
        return torch.abs(self.asarray(a))

    def real(self, a: Any) -> Any:
        a = self.asarray(a)
        return torch.real(a) if a.is_complex() else a

    def imag(self, a: Any) -> Any:
        a = self.asarray(a)
        return torch.imag(a) if a.is_complex() else torch.zeros_like(a)

    def conj(self, a: Any) -> Any:
        return torch.conj(self.asarray(a))

    def diag(self, a: Any) -> Any:
        return torch.diag(self.asarray(a))

    def zeros(self, shape: Any, dtype: Any | None = None) -> Any:
        return torch.zeros(shape, dtype=self._dt(dtype) or torch.float64, device=self.device)

    def ones(self, shape: Any, dtype: Any | None = None) -> Any:
        return torch.ones(shape, dtype=self._dt(dtype) or torch.float64, device=self.device)

    def zeros_like(self, a: Any) -> Any:
        return torch.zeros_like(self.asarray(a))

    def ones_like(self, a: Any) -> Any:
        return torch.ones_like(self.asarray(a))

    def eye(self, n: int, dtype: Any | None = None) -> Any:
        return torch.eye(n, dtype=self._dt(dtype) or torch.float64, device=self.device)

    def kron(self, a: Any, b: Any) -> Any:
        return torch.kron(self.asarray(a), self.asarray(b))

    def square(self, a: Any) -> Any:
        return torch.square(self.asarray(a))

    def copy(self, a: Any) -> Any:
        return self.asarray(a).clone()

    def allclose(self, a: Any, b: Any, rtol: float = 1e-5, atol: float = 1e-8) -> bool:
        return bool(torch.allclose(self.asarray(a), self.asarray(b), rtol=rtol, atol=atol))

    def isclose(self, a: Any, b: Any, rtol: float = 1e-5, atol: float = 1e-8) -> Any:
        return torch.isclose(self.asarray(a), self.asarray(b), rtol=rtol, atol=atol)

    def stack(self, arrays: Any, axis: int = 0) -> Any:
        return torch.stack([self.asarray(x) for x in arrays], dim=axis)

    def concatenate(self, arrays: Any, axis: int = 0) -> Any:
        return torch.cat([self.asarray(x) for x in arrays], dim=axis)

    def arange(self, start: Any, stop: Any | None = None, step: Any = 1) -> Any:
        if stop is None:
            return torch.arange(start, device=self.device)
        return torch.arange(start, stop, step, device=self.device)

    def linspace(self, start: Any, stop: Any, num: int = 50) -> Any:
        return torch.linspace(start, stop, num, dtype=torch.float64, device=self.device)

    def transpose(self, a: Any, axes: Any | None = None) -> Any:
        a = self.asarray(a)
        if axes is None:
            return a.permute(*reversed(range(a.dim())))
        return a.permute(*axes)

    def norm(self, a: Any, ord: Any | None = None, axis: Any | None = None) -> Any:
        return torch.linalg.norm(self.asarray(a), ord=ord, dim=axis)

    def cast(self, a: Any, dtype: Any) -> Any:
        return self.array(a, dtype)

    def sign(self, a: Any) -> Any:
        return torch.sign(self.asarray(a))

    def outer(self, a: Any, b: Any) -> Any:
        return torch.outer(self.asarray(a).reshape(-1), self.asarray(b).reshape(-1))

    # ---- elementary math ----------------------------------------------------------------
    def exp(self, a: Any) -> Any:
        return torch.exp(self.asarray(a))

    def sin(self, a: Any) -> Any:
        return torch.sin(self.asarray(a))

    def cos(self, a: Any) -> Any:
        return torch.cos(self.asarray(a))

    def sqrt(self, a: Any) -> Any:
        return torch.sqrt(self.asarray(a))

    def log(self, a: Any) -> Any:
        return torch.log(self.asarray(a))

    def log2(self, a: Any) -> Any:
        return torch.log2(self.asarray(a))

    # ---- linear algebra -----------------------------------------------------------------
    def svd(self, a: Any, full_matrices: bool = False) -> Tuple[Any, Any, Any]:
        return torch.linalg.svd(self.asarray(a), full_matrices=full_matrices)

    def eigh(self, a: Any) -> Tuple[Any, Any]:
        return torch.linalg.eigh(self.asarray(a))

    def eig(self, a: Any) -> Tuple[Any, Any]:
        return torch.linalg.eig(self.asarray(a))

    def solve(self, a: Any, b: Any, assume_a: str = "gen") -> Any:
        return torch.linalg.solve(self.asarray(a), self.asarray(b))

    def inv(self, a: Any) -> Any:
        return torch.linalg.inv(self.asarray(a))

    def expm(self, a: Any) -> Any:
        return torch.linalg.matrix_exp(self.asarray(a))

    def tensordot(self, a: Any, b: Any, axes: Any = 2) -> Any:
        return torch.tensordot(self.asarray(a), self.asarray(b), dims=axes)

    # ---- random / sampling --------------------------------------------------------------
    def rng(self, seed: int | None = None) -> Any:
        """A numpy Generator: the uniforms stay on the host so that draws are reproducible against the oracle."""
        return np.random.default_rng(seed)

    def normal(self, rng: Any, shape: Tuple[int, ...], dtype: Any | None = None) -> Any:
        return self.array(rng.normal(size=shape), dtype or torch.float64)

    def choice(self, rng: Any, a: int, *, size: int, p: Any | None = None) -> Any:
        """numpy ``Generator.choice(a, size, p=p)`` with its uniforms made explicit (engine.py:415):
        searchsorted(cdf / cdf[-1], rng.random(size), 'right') on the device."""
        u = torch.from_numpy(np.asarray(rng.random(size), dtype=np.float64)).to(self.device)
        if p is None:
            return torch.clamp((u * a).to(torch.int64), max=a - 1)
        cdf = torch.cumsum(self.asarray(p).to(torch.float64).reshape(-1), 0)
        cdf = cdf / cdf[-1]
        return torch.searchsorted(cdf, u, right=True)

    def bincount(self, x: Any, minlength: int = 0) -> Any:
        return torch.bincount(self.asarray(x).to(torch.int64), minlength=minlength)

    def nonzero(self, x: Any) -> Any:
        return torch.nonzero(self.asarray(x)).reshape(-1)

    # ---- autodiff bridge -----------------------------------------------------------------
    def requires_grad(self, x: Any, flag: bool = True) -> Any:
        return x.requires_grad_(flag) if torch.is_tensor(x) else x

    def detach(self, x: Any) -> Any:
        return x.detach() if torch.is_tensor(x) else x

    def vmap(self, fn: Callable[..., Any]) -> Callable[..., Any]:
        """Leading-axis loop (the engine's own batching lives in ``BatchedAnsatz``; op lists are host objects)."""
        def mapped(*args: Any, **kwargs: Any) -> Any:
            nb = next(int(a.shape[0]) for a in args if hasattr(a, "shape") and len(a.shape))
            outs = [fn(*[a[i] if hasattr(a, "shape") and len(a.shape) else a for a in args], **kwargs) for i in range(nb)]
            return torch.stack([self.asarray(o) for o in outs])
        return mapped

    def jit(self, fn: Callable[..., Any]) -> Callable[..., Any]:
        return fn

    def value_and_grad(self, fn: Callable[..., Any], argnums: int | Sequence[int] = 0) -> Callable[..., Any]:
        """(value, gradient) like pytorch_backend.py:446-564: the selected arguments become float64 leaves, the value
        comes back as a Python float (or ndarray), gradients as ndarrays.  ``fn`` typically builds a Circuit from the
        parameters and contracts ``Circuit.state()``: with the engine installed that state is an autograd.Function
        whose backward is the adjoint sweep on the device.  If the tape is broken the reference warns and falls back to
        central differences with eps = 1e-7; so does this (the function evaluations still run on the device)."""
        idx = (argnums,) if isinstance(argnums, int) else tuple(argnums)

        def wrapped(*args: Any, **kwargs: Any) -> Any:
            lst = list(args)
            leaves = []
            for i in idx:
                xi = lst[i]
                ti = (xi.detach().clone() if torch.is_tensor(xi) else torch.tensor(np.asarray(xi, dtype=np.float64)))
                ti = ti.to(self.device).requires_grad_(True)   # device leaves: the state then stays on the device
                lst[i] = ti
                leaves.append(ti)
            try:
                y = fn(*lst, **kwargs)
                if not torch.is_tensor(y):
                    y = torch.as_tensor(y, dtype=torch.float64)
                grads = torch.autograd.grad(y, leaves, allow_unused=True)
                if any(g is None for g in grads):
                    raise RuntimeError("Gradient is None - computation graph may be broken")
                yo = y.detach().cpu().numpy()
                yo = yo.item() if yo.size == 1 else yo
                go = [g.detach().cpu().numpy() for g in grads]
                return yo, (go[0] if len(go) == 1 else tuple(go))
            except Exception as e:  # noqa: BLE001 -- mirrors the reference's behaviour
                warnings.warn(f"autograd failed ({type(e).__name__}: {e}), falling back to finite differences", RuntimeWarning)
                base = [a.detach().cpu().numpy() if torch.is_tensor(a) else a for a in args]
                y0 = float(np.asarray(self.to_numpy(fn(*base, **kwargs)), dtype=np.float64).reshape(-1)[0])
                eps = 1e-7
                out = []
                for i in idx:
                    xi = np.asarray(base[i], dtype=np.float64)
                    g = np.zeros_like(xi)
                    for k in range(xi.size):
                        ap, am = list(base), list(base)
                        xp, xm = xi.copy().reshape(-1), xi.copy().reshape(-1)
                        xp[k] += eps
                        xm[k] -= eps
                        ap[i], am[i] = xp.reshape(xi.shape), xm.reshape(xi.shape)
                        fp = float(np.asarray(self.to_numpy(fn(*ap, **kwargs))).reshape(-1)[0])
                        fm = float(np.asarray(self.to_numpy(fn(*am, **kwargs))).reshape(-1)[0])
                        g.reshape(-1)[k] = (fp - fm) / (2.0 * eps)
                    out.append(g)
                return y0, (out[0] if len(out) == 1 else tuple(out))

        return wrapped


# every name the reference's protocol declares (numerics/api.py:19-121), for the conformance test
PROTOCOL_METHODS = (
    "array asarray to_numpy matmul einsum reshape moveaxis sum mean abs real imag conj diag zeros ones zeros_like ones_like "
    "eye kron square stack concatenate arange linspace transpose norm cast sign outer exp sin cos sqrt log log2 choice "
    "bincount nonzero svd eigh eig solve inv expm tensordot rng normal requires_grad detach").split()
PROTOCOL_ATTRS = "name complex64 complex128 float32 float64 int32 int64 int8 bool int".split()
