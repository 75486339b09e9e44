"""Pauli-sum Hamiltonians in bitmask form for the matrix-free device kernels.

Replaces the dense 2^n x 2^n construction of the reference
(libs/quantum_library/kernels/pauli.py:65-87, ``pauli_string_sum_dense``) and the densified
sparse-H mat-vec of ``apply_op`` (applications/chem/chem_libs/hamiltonians_chem_library/
hamiltonian_builders.py:283-318).  A term  c * P  with P a Pauli string is stored as
(xmask, zmask, c * i^{#Y}) so that  P|j> = i^{#Y} (-1)^{popc(j & zmask)} |j ^ xmask>.
Qubit q of an n-qubit register is index bit n-1-q (big-endian, statevector.py:28-42).
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, Iterable, List, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .program import _prep

_CODE = {"I": 0, "X": 1, "Y": 2, "Z": 3, 0: 0, 1: 1, 2: 2, 3: 3}


class PauliSum:
    """sum_t coef_t * P_t on n qubits, grouped by xmask for the device kernels."""

    def __init__(self, n: int, terms: Iterable[Tuple[int, int, complex]]) -> None:
        """terms: iterable of (xmask, zmask, coef) with the i^{#Y} phase already folded in."""
        self.n = int(n)
        acc: Dict[Tuple[int, int], complex] = OrderedDict()
        for x, z, c in terms:
            key = (int(x), int(z))
            acc[key] = acc.get(key, 0.0) + complex(c)
        groups: Dict[int, List[Tuple[int, complex]]] = OrderedDict()
        for (x, z), c in acc.items():
            if c != 0:
                groups.setdefault(x, []).append((z, c))
        self.group_x = np.array(list(groups.keys()), dtype=np.uint64)
        ptr = [0]
        zs: List[int] = []
        cs: List[complex] = []
        for x, lst in groups.items():
            for z, c in lst:
                zs.append(z)
                cs.append(c)
            ptr.append(len(zs))
        self.group_ptr = np.array(ptr, dtype=np.int32)
        self.term_z = np.array(zs, dtype=np.uint64)
        self.term_coef = np.array(cs, dtype=np.complex128)
        self._dev: Dict[Any, Tuple[torch.Tensor, ...]] = {}

    # ---- constructors -------------------------------------------------------------------
    @classmethod
    def from_codes(cls, terms: Sequence[Sequence[int]], weights: Sequence[complex] | None = None) -> "PauliSum":
        """Reference list format (kernels/pauli.py:74-87): each term a length-n list of codes
        0,1,2,3 = I,X,Y,Z, list position = qubit."""
        n = len(terms[0])
        w = [1.0] * len(terms) if weights is None else list(weights)
        out = []
        for ps, c in zip(terms, w):
            out.append(cls._term(n, [(q, int(code)) for q, code in enumerate(ps) if int(code)], c))
        return cls(n, out)

    @classmethod
    def from_pauli_list(cls, n: int, ham: Sequence[Tuple[complex, Sequence[Tuple[str, int]]]]) -> "PauliSum":
        """[(coeff, [(P, q), ...]), ...] (libs/hamiltonian_encoding/hamiltonian_grouping.py:120-138)."""
        return cls(n, [cls._term(n, [(int(q), _CODE[str(p).upper()]) for (p, q) in ops], c) for c, ops in ham])

    @classmethod
    def from_qubit_operator(cls, n: int, qop: Any) -> "PauliSum":
        """Anything with an OpenFermion-style ``terms`` dict {((q, 'X'), ...): coeff}."""
        return cls(n, [cls._term(n, [(int(q), _CODE[str(p).upper()]) for (q, p) in t], c) for t, c in qop.terms.items()])

    @staticmethod
    def _term(n: int, ops: Sequence[Tuple[int, int]], coef: complex) -> Tuple[int, int, complex]:
        x = z = ny = 0
        for q, code in ops:
            bit = 1 << (n - 1 - q)
            if code in (1, 2):
                x |= bit
            if code in (2, 3):
                z |= bit
            if code == 2:
                ny += 1
        return x, z, complex(coef) * (1j ** ny)

    # ---- device side --------------------------------------------------------------------
    @property
    def n_terms(self) -> int:
        return int(self.term_z.size)

    @property
    def n_groups(self) -> int:
        return int(self.group_x.size)

    def is_diagonal(self) -> bool:
        return self.n_groups == 0 or (self.n_groups == 1 and int(self.group_x[0]) == 0)

    def to_device(self, device: torch.device) -> Tuple[torch.Tensor, ...]:
        key = str(device)
        if key not in self._dev:
            def up(a: np.ndarray, view=None) -> torch.Tensor:
                a = np.ascontiguousarray(a)
                if view is not None:
                    a = a.view(view)
                if a.size == 0:
                    a = np.zeros(1, dtype=a.dtype)
                return torch.from_numpy(a.copy()).to(device)
            self._dev[key] = (up(self.group_x, np.int64), up(self.group_ptr), up(self.term_z, np.int64),
                              up(self.term_coef.view(np.float64)))
        return self._dev[key]

    def expectation(self, state: torch.Tensor, *, global_base: int = 0) -> torch.Tensor:
        """<psi|H|psi> per batch member (complex128 tensor [batch]); one read pass per xmask group."""
        ptr, n, batch, dt, stream = _prep(state)
        if n != self.n and global_base == 0:
            raise _lib.TqbError(f"PauliSum on {self.n} qubits applied to a {n}-qubit state")
        gx, gp, tz, tc = self.to_device(state.device)
        out = torch.empty((batch, 2), dtype=torch.float64, device=state.device)
        with torch.cuda.device(state.device):
            _lib.check(_lib.load().tqb_expect_pauli_sum(ptr, n, batch, dt, global_base, gx.data_ptr(), gp.data_ptr(),
                                                        self.n_groups, tz.data_ptr(), tc.data_ptr(), out.data_ptr(), stream))
        return torch.view_as_complex(out)

    def apply(self, state: torch.Tensor, out: torch.Tensor | None = None, *, global_base: int = 0) -> torch.Tensor:
        """out = H |psi> (replaces apply_op)."""
        ptr, n, batch, dt, stream = _prep(state)
        if out is None:
            out = torch.empty_like(state)
        gx, gp, tz, tc = self.to_device(state.device)
        with torch.cuda.device(state.device):
            _lib.check(_lib.load().tqb_apply_pauli_sum(ptr, out.data_ptr(), n, batch, dt, global_base, gx.data_ptr(),
                                                       gp.data_ptr(), self.n_groups, tz.data_ptr(), tc.data_ptr(), stream))
        return out
