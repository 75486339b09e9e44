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
        self._tiled: Dict[Any, Any] = {}
        # Hermitian groups (real Pauli coefficients: stored coef / i^#Y real) let the tiled kernel visit each pair
        # (j, j ^ x) once; purely real stored coefficients let it skip the imaginary phase sums
        ny = np.array([bin(int(x) & int(z)).count("1") for x, z in self._term_xz()], dtype=np.int64) if len(zs) else np.zeros(0, np.int64)
        orig = self.term_coef / (1j ** ny) if len(zs) else self.term_coef
        scale = float(np.abs(self.term_coef).max()) if len(zs) else 1.0
        self.hermitian = bool(np.all(np.abs(orig.imag) <= 1e-14 * max(scale, 1e-300)))
        self.real_coef = bool(np.all(np.abs(self.term_coef.imag) <= 1e-14 * max(scale, 1e-300)))

    def _term_xz(self):
        for g in range(len(self.group_x)):
            for t in range(int(self.group_ptr[g]), int(self.group_ptr[g + 1])):
                yield int(self.group_x[g]), int(self.term_z[t])

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

    # ---- tile layouts for the staged kernel ---------------------------------------------------------------
    @staticmethod
    def plan_layouts(group_x: Sequence[int], n: int, m: int, l_min: int) -> List[Tuple[List[int], List[int]]] | None:
        """Sort the xmask groups into tile layouts: [(tile bits ascending, [group indices]), ...]; every group's xmask lies
        inside its layout's tile bits, which always hold the ``l_min`` lowest index bits (contiguous runs for coalesced
        loads) and ``m`` bits in all.  Greedy: a layout keeps taking the group that needs the fewest new bits.  None when
        an xmask does not fit any tile (more than m - l_min bits outside the low ones)."""
        m = min(m, n)
        l_min = min(l_min, m)
        low = (1 << l_min) - 1
        todo = [g for g in range(len(group_x)) if int(group_x[g]) & ~low]
        easy = [g for g in range(len(group_x)) if not (int(group_x[g]) & ~low)]   # x == 0 and masks on the low bits: any layout
        layouts: List[Tuple[List[int], List[int]]] = []
        while todo or (easy and not layouts):
            bits = low
            chosen: List[int] = []
            while True:
                best, best_need = None, None
                for g in todo:
                    need = int(group_x[g]) & ~bits
                    cnt = bin(need).count("1")
                    if bin(bits).count("1") + cnt <= m and (best is None or cnt < best_need[0] or (cnt == best_need[0] and need < best_need[1])):
                        best, best_need = g, (cnt, need)
                if best is None:
                    break
                bits |= int(group_x[best])
                chosen.append(best)
                todo.remove(best)
            if not chosen and todo:
                return None
            b = 0
            while bin(bits).count("1") < m and b < n:   # fill up with the lowest free bits (longer contiguous runs)
                bits |= 1 << b
                b += 1
            layouts.append(([p for p in range(n) if (bits >> p) & 1], chosen))
        if layouts:
            layouts[0] = (layouts[0][0], easy + layouts[0][1])
        return layouts

    def _tiled_plan(self, n: int, itemsize: int, device: torch.device):
        key = (n, itemsize, str(device))
        if key in self._tiled:
            return self._tiled[key]
        plan = None
        if self.n_groups and n >= 6:
            big = 13 if itemsize == 16 else 14
            m = n if n <= big else (12 if itemsize == 16 else 13)   # 64 KiB tiles when the state streams
            m = min(m, n)
            lay = self.plan_layouts([int(x) for x in self.group_x], n, m, 3 if itemsize == 16 else 4)
            if lay is not None and len(lay) <= 64:
                import ctypes as C
                arr = (_lib.PauliLayout * len(lay))()
                gxl: List[int] = []
                gptr = [0]
                zl: List[int] = []
                zout: List[int] = []
                coef: List[complex] = []
                ok = True
                for li, (bits, groups) in enumerate(lay):
                    L = 0
                    while L < len(bits) and bits[L] == L:
                        L += 1
                    mask = sum(1 << b for b in bits)
                    pos = {b: k for k, b in enumerate(bits)}
                    arr[li].m, arr[li].L = len(bits), L
                    arr[li].group_begin, arr[li].n_groups = len(gxl), len(groups)
                    for j, b in enumerate(bits[L:]):
                        arr[li].hb[j] = b
                    nt0 = len(zl)
                    for g in groups:
                        x = int(self.group_x[g])
                        gxl.append(sum(1 << pos[b] for b in range(n) if (x >> b) & 1))
                        for t in range(int(self.group_ptr[g]), int(self.group_ptr[g + 1])):
                            z = int(self.term_z[t])
                            zl.append(sum(1 << pos[b] for b in bits if (z >> b) & 1))
                            zout.append(z & ~mask)
                            coef.append(complex(self.term_coef[t]))
                        gptr.append(len(zl))
                    arr[li].n_terms = len(zl) - nt0
                    if arr[li].n_terms * 20 + (itemsize << len(bits)) + (8 << (len(bits) - L)) + 64 > 200 * 1024:
                        ok = False
                if ok:
                    plan = (arr, len(lay), torch.from_numpy(np.asarray(gxl, dtype=np.uint32).view(np.int32).copy()).to(device),
                            torch.from_numpy(np.asarray(gptr, dtype=np.int32)).to(device),
                            torch.from_numpy(np.asarray(zl, dtype=np.uint32).view(np.int32).copy()).to(device),
                            torch.from_numpy(np.asarray(zout, dtype=np.uint64).view(np.int64).copy()).to(device),
                            torch.from_numpy(np.asarray(coef, dtype=np.complex128).view(np.float64).copy()).to(device))
        self._tiled[key] = plan
        return plan

    def expectation(self, state: torch.Tensor, *, global_base: int = 0, tiled: bool | None = None) -> torch.Tensor:
        """<psi|H|psi> per batch member (complex128 tensor [batch]).  Tile-staged evaluation (one read of the state per
        tile layout, csrc/tqb_reduce.cu expect_pauli_tiled_kernel) whenever the xmasks fit tile layouts; ``tiled=False``
        forces the gather kernel (one global gather per xmask group and amplitude)."""
        ptr, n, batch, dt, stream = _prep(state)
        if n != self.n and global_base == 0:
            raise _lib.TqbError(f"PauliSum on {self.n} qubits applied to a {n}-qubit state")
        out = torch.empty((batch, 2), dtype=torch.float64, device=state.device)
        plan = self._tiled_plan(n, state.element_size(), state.device) if tiled is not False else None
        if plan is not None and batch <= 65535:
            import ctypes as C
            arr, nl, gxl, gptr, zl, zout, coef = plan
            flags = (1 if self.hermitian else 0) | (2 if self.real_coef else 0)
            with torch.cuda.device(state.device):
                _lib.check(_lib.load().tqb_expect_pauli_tiled(ptr, n, batch, dt, global_base, C.cast(arr, C.c_void_p), nl,
                                                              gxl.data_ptr(), gptr.data_ptr(), zl.data_ptr(), zout.data_ptr(),
                                                              coef.data_ptr(), flags, out.data_ptr(), stream))
            return torch.view_as_complex(out)
        if tiled:
            raise _lib.TqbError("this Pauli sum does not fit tile layouts")
        gx, gp, tz, tc = self.to_device(state.device)
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
