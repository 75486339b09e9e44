"""UCC statevector energy + adjoint gradient on the device (the H2O-UCCSD config).

Host-side mirror of the reference's numeric UCC path, restricted to what the hot loop needs:

  get_statevector / evolve_excitation   applications/chem/chem_libs/quantum_chem_library/statevector_ops.py:68-168
  energy_and_grad_statevector           statevector_ops.py:203-244
  get_init_circuit (HF occupation)      statevector_ops.py:172-199
  UCCSD excitation enumeration          applications/chem/algorithms/ucc.py:680-832, uccsd.py:268-318
  get_hop_from_integral                 chem_libs/hamiltonians_chem_library/hamiltonian_builders.py:71-108

Differences in mechanism (not in results):
  * exp(theta G) for G = a+_p a_q - h.c. (or the double) is a Givens rotation between two basis
    patterns of the 2 (4) target qubits, signed by the Jordan-Wigner Z-string parity: ONE PAIR
    gate touching 2/2^k of the amplitudes, instead of two dense k-qubit mat-vecs + a sign vector
    + an axpy (statevector_ops.py:152-168).
  * H|psi> is a matrix-free Pauli sum (Jordan-Wigner done here with bitmasks; OpenFermion is
    not needed), instead of a densified 2^n x 2^n matrix (hamiltonian_builders.py:296-299).
  * The gradient is an analytic adjoint sweep (model: civector_ops.py:141-200) instead of P+1
    finite-difference evaluations (numerics/backends/numpy_backend.py:386-454).
Orbital k <-> index bit k <-> TyxonQ qubit n-1-k (statevector_ops.py:34).
"""
from __future__ import annotations

from itertools import product
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from . import program as P
from .gates import LGate, PAIR, pair_gate
from .pauli import PauliSum
from .planner import TileConfig, compile_program, default_tile

XZ = Tuple[int, int]


# ---- Jordan-Wigner in symplectic (xmask, zmask) form:  term = coef * X^x Z^z ----------------
def _ladder(k: int, dagger: bool) -> Dict[XZ, complex]:
    """a_k = X_k (1 - Z_k)/2 * Z_{<k},  a+_k = X_k (1 + Z_k)/2 * Z_{<k}."""
    below = (1 << k) - 1
    x = 1 << k
    return {(x, below): 0.5, (x, below | x): 0.5 if dagger else -0.5}


def _mul(A: Dict[XZ, complex], B: Dict[XZ, complex]) -> Dict[XZ, complex]:
    out: Dict[XZ, complex] = {}
    for (x1, z1), c1 in A.items():
        for (x2, z2), c2 in B.items():
            s = -1.0 if bin(z1 & x2).count("1") & 1 else 1.0
            key = (x1 ^ x2, z1 ^ z2)
            out[key] = out.get(key, 0.0) + s * c1 * c2
    return out


def jw_word(word: Sequence[Tuple[int, int]]) -> Dict[XZ, complex]:
    """JW image of a product of ladder operators [(orbital, dagger), ...] in operator order."""
    acc: Dict[XZ, complex] = {(0, 0): 1.0}
    for k, dg in word:
        acc = _mul(acc, _ladder(int(k), bool(dg)))
    return {k: v for k, v in acc.items() if abs(v) > 1e-15}


def _letters_key(x: int, z: int, n: int) -> Tuple[Tuple[int, str], ...]:
    out = []
    for q in range(n):
        xb, zb = (x >> q) & 1, (z >> q) & 1
        if xb or zb:
            out.append((q, "X" if (xb and not zb) else ("Y" if xb else "Z")))
    return tuple(out)


def excitation_zmask_sign(f_idx: Sequence[int], n: int) -> Tuple[int, int]:
    """Z-string mask and global sign of G = T - T^+ exactly as the reference derives them
    (statevector_ops.py:45-54): Z positions of a JW term of T (all terms share them), and
    sign = +1 iff the coefficient of the lexicographically smallest term has positive real part.
    In letter form X^x Z^z carries (-i)^{#Y}; the smallest term is all-X on f_idx (#Y = 0)."""
    k = len(f_idx)
    word = [(int(f_idx[i]), 1 if i < k // 2 else 0) for i in range(k)]
    terms = jw_word(word)
    fmask = 0
    for i in f_idx:
        fmask |= 1 << int(i)
    first = min(terms.keys(), key=lambda xz: _letters_key(xz[0], xz[1], n))
    x, z = first
    ny = bin(x & z).count("1")
    coef = terms[first] * ((-1j) ** ny)
    zmask = z & ~fmask
    return zmask, (1 if coef.real > 0 else -1)


def excitation_gate(f_idx: Sequence[int], theta: float, n: int, *, mode: str = "fermion", param: Optional[int] = None) -> LGate:
    """exp(theta * G) as a PAIR gate.  Listed order (p, q[, r, s]) = most significant matrix bit first
    (statevector_ops.py:34-40): T maps pattern A = |0..01..1> (annihilated orbitals occupied) to
    B = |1..10..0>, so on (A, B): exp(+-theta G) = [[c, -+s], [+-s, c]]."""
    k = len(f_idx)
    half = k // 2
    pat_a = (1 << half) - 1
    pat_b = pat_a << half
    c, s = float(np.cos(theta)), float(np.sin(theta))
    zmask, sign = (0, 1)
    if mode == "fermion":
        zmask, sign = excitation_zmask_sign(f_idx, n)
    s *= sign
    even = np.array([[c, -s], [s, c]], dtype=np.complex128)
    odd = np.array([[c, s], [-s, c]], dtype=np.complex128)
    qubits = [n - 1 - int(i) for i in f_idx]
    g = pair_gate(even, qubits, n, pat_a, pat_b, m2_odd=odd, zmask=zmask, name="ucc", param=param)
    g.sign = sign  # type: ignore[attr-defined]
    return g


def hf_basis_index(n: int, n_elec_s: Tuple[int, int]) -> int:
    """statevector_ops.py:190-194: X on wires n-1-i (i < nb) and n/2-1-i (i < na) = orbitals i and n/2+i."""
    na, nb = int(n_elec_s[0]), int(n_elec_s[1])
    idx = 0
    for i in range(nb):
        idx |= 1 << i
    for i in range(na):
        idx |= 1 << (n // 2 + i)
    return idx


def uccsd_ex_ops(no: int, nv: int) -> Tuple[List[tuple], List[int]]:
    """ucc.py:680-832 with init_method="zeros" (no screening/sorting, uccsd.py:257-261).  Spin-orbital
    numbering: beta-occ i, beta-virt no+a, alpha-occ no+nv+i, alpha-virt 2no+nv+a (ucc.py:768-778)."""
    def ao(i): return no + nv + i
    def av(a): return 2 * no + nv + a
    def bo(i): return i
    def bv(a): return no + a
    ex1, id1, pid = [], [], -1
    for i in range(no):
        for a in range(nv):
            pid += 1
            ex1 += [(av(a), ao(i)), (bv(a), bo(i))]
            id1 += [pid, pid]
    ex2, id2, pid = [], [], -1
    for i in range(no):
        for j in range(i):
            for a in range(nv):
                for b in range(a):
                    pid += 1
                    ex2 += [(av(b), av(a), ao(i), ao(j)), (bv(b), bv(a), bo(i), bo(j))]
                    id2 += [pid, pid]
    for i in range(no):
        for j in range(i + 1):
            for a in range(nv):
                for b in range(a + 1):
                    pid += 1
                    if i == j and a == b:
                        ex2.append((bv(a), av(a), ao(i), bo(i)))
                        id2.append(pid)
                        continue
                    ex2 += [(bv(b), av(a), ao(i), bo(j)), (av(b), bv(a), bo(i), ao(j))]
                    id2 += [pid, pid]
                    if i != j and a != b:
                        pid += 1
                        ex2 += [(bv(a), av(b), ao(i), bo(j)), (av(a), bv(b), bo(i), ao(j))]
                        id2 += [pid, pid]
    off = max(id1) + 1 if id1 else 0
    return ex1 + ex2, id1 + [i + off for i in id2]


def random_integral(nao: int, seed: int = 2077) -> Tuple[np.ndarray, np.ndarray]:
    """hamiltonian_builders.py:261-278 (the reference's own synthetic-integral generator)."""
    np.random.seed(seed)
    int1e = np.random.uniform(-1, 1, size=(nao, nao))
    int2e = np.random.uniform(-1, 1, size=(nao, nao, nao, nao))
    int1e = 0.5 * (int1e + int1e.T)
    int2e = 0.25 * (int2e + int2e.transpose((0, 1, 3, 2)) + int2e.transpose((1, 0, 2, 3)) + int2e.transpose((2, 3, 0, 1)))
    int2e = 0.5 * (int2e + int2e.transpose(3, 2, 1, 0))
    return int1e, int2e


def hamiltonian_from_integral(int1e: np.ndarray, int2e: np.ndarray) -> PauliSum:
    """hamiltonian_builders.py:71-108 followed by Jordan-Wigner, as a PauliSum over index bits."""
    n_orb = int1e.shape[0]
    ns = 2 * n_orb
    acc: Dict[XZ, complex] = {}

    def add(word, v):
        for key, c in jw_word(word).items():
            acc[key] = acc.get(key, 0.0) + v * c

    for p, q in product(range(ns), repeat=2):
        if (p < n_orb) == (q < n_orb):
            v = int1e[p % n_orb, q % n_orb]
            if abs(v) >= 1e-12:
                add([(p, 1), (q, 0)], v)

    def h2(p, q, r, s):
        if ((p < n_orb) == (s < n_orb)) and ((q < n_orb) == (r < n_orb)):
            return int2e[p % n_orb, s % n_orb, q % n_orb, r % n_orb]
        return 0.0

    for q, s in product(range(ns), repeat=2):
        for p, r in product(range(q), range(s)):
            v = h2(p, q, r, s) - h2(q, p, r, s)
            if abs(v) >= 1e-12:
                add([(p, 1), (q, 1), (r, 0), (s, 0)], v)
    # X^x Z^z = (-i)^{#Y} * letters  ->  PauliSum wants coef * i^{#Y} * X^x Z^z with `coef` the letter coefficient:
    # coef_xz * X^x Z^z is already in the kernel's form (P|j> = coef (-1)^{popc(j&z)} |j^x>).
    terms = [(x, z, c) for (x, z), c in acc.items() if abs(c) > 1e-14]
    return PauliSum(ns, terms)


class UCCStatevector:
    """Energy and adjoint gradient of a UCC ansatz on one device (complex128 by default)."""

    def __init__(self, n: int, n_elec_s: Tuple[int, int], ex_ops: Sequence[tuple], param_ids: Sequence[int],
                 hamiltonian: PauliSum, *, mode: str = "fermion", device: str | torch.device = "cuda",
                 dtype: torch.dtype = torch.complex128, tile: Optional[TileConfig] = None) -> None:
        self.n = int(n)
        self.n_elec_s = tuple(n_elec_s)
        self.ex_ops = [tuple(e) for e in ex_ops]
        self.param_ids = [int(i) for i in param_ids]
        self.n_params = max(self.param_ids) + 1 if self.param_ids else 0
        self.ham = hamiltonian
        self.mode = mode
        self.device = torch.device(device)
        self.dtype = dtype
        itemsize = 16 if dtype == torch.complex128 else 8
        self.tile = tile or default_tile(self.n, itemsize, 2)
        self.hf_index = hf_basis_index(self.n, self.n_elec_s)
        _lib.ensure_device(self.device.index or 0)
        # structure (zmask, sign, patterns) is parameter independent: lower once with theta = 0
        self._proto = [excitation_gate(f, 0.0, self.n, mode=mode, param=pid) for f, pid in zip(self.ex_ops, self.param_ids)]
        self._signs = np.array([g.sign for g in self._proto], dtype=np.float64)  # type: ignore[attr-defined]
        self._fwd = compile_program(self._proto, self.n, self.tile, itemsize=itemsize)
        one = TileConfig(m=self.tile.m, L=self.tile.L, threads=self.tile.threads, ctas_per_sm=self.tile.ctas_per_sm, max_gates=1)
        self._rev = compile_program(list(reversed(self._proto)), self.n, one, itemsize=itemsize)
        assert self._rev.order == list(range(len(self._proto)))
        self._fwd_dev = P.DeviceProgram(self._fwd, self.device, dtype)
        self._rev_dev = P.DeviceProgram(self._rev, self.device, dtype)
        # whole-state PAIR descriptors for the gradient reductions (m = n: local bit == index bit)
        full = compile_program(list(reversed(self._proto)), self.n, TileConfig(m=self.n, L=min(self.tile.L, self.n), max_gates=1),
                               itemsize=itemsize)
        self._grad_descs = np.ascontiguousarray(full.gates)
        self._kb = torch.empty((2, 1 << self.n), dtype=dtype, device=self.device)
        self._gout = torch.zeros(max(self.n_params, 1), dtype=torch.float64, device=self.device)
        # small states: the whole forward / reverse sweep is ONE persistent launch each (tqb_pair_sweep)
        self.use_sweep = self.n <= 26 and len(self._proto) > 0
        N = len(self._proto)
        self._steps_np = np.zeros(2 * N, dtype=_lib.PAIR_STEP_DTYPE)   # [forward steps | reverse steps]
        fwd_full = compile_program(self._proto, self.n, TileConfig(m=self.n, L=min(self.tile.L, self.n), max_gates=1), itemsize=itemsize)
        assert fwd_full.order == list(range(N))
        for half, descs in ((0, fwd_full.gates), (1, self._grad_descs)):
            for i in range(N):
                e = self._steps_np[half * N + i]
                d = descs[i]
                j = i if half == 0 else N - 1 - i
                e["k"] = d["k"]
                e["slot"] = self.param_ids[j]
                e["sbits"][:] = d["sbits"]
                e["off_a"], e["off_b"], e["zmask"] = d["off_a"], d["off_b"], d["zmask"]
                e["scale"] = -2.0 * float(self._signs[j])
        self._steps_host = torch.zeros(self._steps_np.nbytes, dtype=torch.uint8).pin_memory()
        self._steps_dev = torch.zeros(max(self._steps_np.nbytes, 8), dtype=torch.uint8, device=self.device)
        self._sync = torch.zeros(2, dtype=torch.int64, device=self.device)

    # -- matrices for the current parameters -------------------------------------------------
    def _fill(self, prog, params: np.ndarray, reverse: bool) -> np.ndarray:
        mats = np.empty(prog.mats.size, dtype=np.complex128)
        N = len(self._proto)
        idx = np.asarray(prog.order)
        src = (N - 1 - idx) if reverse else idx          # index into ex_ops for the i-th scheduled gate
        th = params[np.asarray(self.param_ids)[src]]
        sg = self._signs[src]
        c = np.cos(th)
        s = np.sin(th) * sg * (-1.0 if reverse else 1.0)  # dagger = rotation by -theta
        blk = np.stack([c, -s, s, c, c, s, -s, c], axis=1).astype(np.complex128)  # even | odd
        has_par = prog.gates["zmask"] != 0
        off = prog.gates["mat_off"].astype(np.int64)
        for i in range(N):  # gates without parity store only the even block
            w = 8 if has_par[i] else 4
            mats[off[i]:off[i] + w] = blk[i, :w]
        return mats

    def statevector(self, params: Sequence[float]) -> torch.Tensor:
        """get_statevector (statevector_ops.py:68-116): HF state evolved by every excitation; device tensor."""
        params = np.asarray(params, dtype=np.float64)
        state = P.new_state(self.n, dtype=self.dtype, device=self.device, basis_index=self.hf_index)
        if self._proto:
            self._fwd_dev.upload_mats(self._fill(self._fwd, params, False))
            self._fwd_dev.run(state)
        return state

    def energy(self, params: Sequence[float]) -> float:
        psi = self.statevector(params)
        return float(self.ham.expectation(psi)[0].real.cpu())

    def _enqueue(self) -> None:
        """Everything of one energy + gradient evaluation on the current stream, device side only: H2D of the
        two matrix buffers, HF state, forward passes, |bra> = H|ket>, <ket|bra>, and the reverse sweep.  No
        host synchronisation and no allocation, so the whole thing is captured into one CUDA graph."""
        lib = _lib.load()
        kb = self._kb
        ptr, n, _, dt, stream = P._prep(kb[0])
        _lib.check(lib.tqb_init_basis(ptr, n, 1, dt, 0, self.hf_index, stream))
        N = len(self._proto)
        if self.use_sweep:
            self._steps_dev.copy_(self._steps_host, non_blocking=True)
            _lib.check(lib.tqb_pair_sweep(kb[0].data_ptr(), 0, n, dt, self._steps_dev.data_ptr(), N, 0, 0,
                                          self._sync.data_ptr(), stream))
        elif self._proto:
            self._fwd_dev.upload()
            self._rev_dev.upload()
            self._fwd_dev.run(kb[0])
        self.ham.apply(kb[0], kb[1])
        _lib.check(lib.tqb_inner(kb[0].data_ptr(), kb[1].data_ptr(), n, 1, dt, self._e.data_ptr(), stream))
        self._gout.zero_()
        ket_ptr, bra_ptr = kb[0].data_ptr(), kb[1].data_ptr()
        if self.use_sweep:
            step_bytes = self._steps_np.dtype.itemsize
            _lib.check(lib.tqb_pair_sweep(ket_ptr, bra_ptr, n, dt, self._steps_dev.data_ptr() + N * step_bytes, N, 1,
                                          self._gout.data_ptr(), self._sync.data_ptr() + 8, stream))
            N = 0  # the per-gate loop below is the large-state path
        t = self.tile
        passes = self._rev_dev._passes
        psz = passes.dtype.itemsize
        gsz = self._grad_descs.dtype.itemsize
        for i in range(N):
            j = N - 1 - i
            _lib.check(lib.tqb_grad_pair(bra_ptr, ket_ptr, n, dt, self._grad_descs.ctypes.data + i * gsz,
                                         -2.0 * float(self._signs[j]), self._gout.data_ptr(), self.param_ids[j], stream))
            if i + 1 < N:  # the last un-apply is not needed
                _lib.check(lib.tqb_run_passes(kb.data_ptr(), n, 2, dt, 0, passes.ctypes.data + i * psz, 1,
                                              self._rev_dev.gates_dev.data_ptr(), self._rev_dev.mats_dev.data_ptr(),
                                              t.threads, t.ctas_per_sm, stream))
        self._out[0:1].copy_(self._e[0:1])
        self._out[1:1 + self.n_params].copy_(self._gout[: self.n_params])

    def energy_and_grad(self, params: Sequence[float], *, graph: bool = True, _sync: bool = True) -> Tuple[float, np.ndarray]:
        """energy_and_grad_statevector (statevector_ops.py:203-244) with an analytic adjoint gradient.
        ``graph``: replay the evaluation as one CUDA graph (captured on the second call) instead of ~300
        separate launches; new parameters only rewrite the pinned matrix buffers."""
        params = np.asarray(params, dtype=np.float64)
        if not hasattr(self, "_e"):
            self._e = torch.zeros(2, dtype=torch.float64, device=self.device)
            self._out = torch.zeros(1 + max(self.n_params, 1), dtype=torch.float64, device=self.device)
            self._out_host = torch.zeros(1 + max(self.n_params, 1), dtype=torch.float64).pin_memory()
            self._graph = None
            self._calls = 0
        if self._proto and self.use_sweep:
            N = len(self._proto)
            th = params[np.asarray(self.param_ids)]
            c, sn = np.cos(th), np.sin(th) * self._signs
            st = self._steps_np
            st["c"][:N], st["s"][:N] = c, sn                      # forward: rotation by +theta (sign folded in)
            st["c"][N:], st["s"][N:] = c[::-1], -sn[::-1]         # reverse: un-apply, last excitation first
            self._steps_host.copy_(torch.from_numpy(st.view(np.uint8).reshape(-1)))
        elif self._proto:
            self._fwd_dev.fill_host(self._fill(self._fwd, params, False))
            self._rev_dev.fill_host(self._fill(self._rev, params, True))
        slot, n_slots = getattr(self, "_ws_slot", (0, 1))
        lib = _lib.load()
        with torch.cuda.device(self.device):
            # concurrent replicas (energy_and_grad_batch) reduce into their own slices of the library's workspace
            _lib.check(lib.tqb_workspace_slot(slot, n_slots))
            try:
                if graph and self._graph is None and self._calls >= 1:
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._enqueue()
                    self._graph = g
                if graph and self._graph is not None:
                    self._graph.replay()
                else:
                    self._enqueue()
            finally:
                lib.tqb_workspace_slot(0, 1)
            self._calls += 1
            self._out_host.copy_(self._out, non_blocking=True)
            if not _sync:
                return None  # type: ignore[return-value]   (energy_and_grad_batch collects after every replica is launched)
            torch.cuda.current_stream().synchronize()
        out = self._out_host.numpy()
        return float(out[0]), out[1:1 + self.n_params].copy()

    MAX_REPLICAS = 32   # slices of the reduction workspace (tqb_workspace_slot)

    def energy_and_grad_batch(self, params: np.ndarray, *, replicas: int = 32) -> Tuple[np.ndarray, np.ndarray]:
        """Many parameter vectors ([B, n_params] -> energies [B], gradients [B, n_params]).  An evaluation of a small
        molecule is a chain of ~300 dependent steps (latency, not throughput: the sweeps of a register of <= 15 qubits run in ONE
        CTA whose barrier is __syncthreads, pair_sweep_cta_kernel), so ``replicas`` independent
        evaluations run CONCURRENTLY: each replica has its own state buffers, stream and CUDA graph, the persistent sweep
        kernels of different replicas share the device, and every replica reduces into its own slice of the
        library's workspace (tqb_workspace_slot: the partial sums of <ket|bra> must not be shared between streams)."""
        p = np.asarray(params, dtype=np.float64).reshape(-1, max(self.n_params, 1))
        B = p.shape[0]
        R = max(1, min(int(replicas), B, self.MAX_REPLICAS))
        if not hasattr(self, "_replicas") or len(self._replicas) < R:
            self._replicas = [self] + [UCCStatevector(self.n, self.n_elec_s, self.ex_ops, self.param_ids, self.ham, mode=self.mode,
                                                     device=self.device, dtype=self.dtype, tile=self.tile) for _ in range(R - 1)]
            self._streams = [torch.cuda.Stream(device=self.device) for _ in range(R)]
            for r, rep in enumerate(self._replicas):
                if getattr(rep, "_ws_slot", None) != (r, self.MAX_REPLICAS):   # (graphs bake the slice in: re-capture on change)
                    rep._ws_slot = (r, self.MAX_REPLICAS)
                    rep._graph = None
                    if hasattr(rep, "_calls"):
                        rep._calls = 0
            for r, (rep, st) in enumerate(zip(self._replicas, self._streams)):   # warm up + capture each replica's graph on ITS stream
                with torch.cuda.stream(st):
                    rep.energy_and_grad(p[0])
                    rep.energy_and_grad(p[0])
        es = np.empty(B)
        gs = np.empty((B, self.n_params))
        with torch.cuda.device(self.device):
            for b0 in range(0, B, R):
                live = []
                for r in range(min(R, B - b0)):
                    with torch.cuda.stream(self._streams[r]):
                        self._replicas[r].energy_and_grad(p[b0 + r], _sync=False)
                    live.append(r)
                for r in live:
                    self._streams[r].synchronize()
                    out = self._replicas[r]._out_host.numpy()
                    es[b0 + r] = out[0]
                    gs[b0 + r] = out[1:1 + self.n_params]
        return es, gs
