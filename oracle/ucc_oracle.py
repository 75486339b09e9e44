"""CPU oracle for the UCC statevector path (H2O-UCCSD config)  --  TEST INFRASTRUCTURE ONLY.

Restates, in numpy/scipy and WITHOUT OpenFermion/PySCF (neither is installed here):

* ``evolve_excitation``            applications/chem/chem_libs/quantum_chem_library/statevector_ops.py:140-168
* ``apply_excitation_statevector`` statevector_ops.py:25-63  (local matrix + JW Z-string sign vector)
* ``get_init_circuit`` (HF state)  statevector_ops.py:172-199
* ``energy_and_grad_statevector``  statevector_ops.py:203-244
* UCCSD excitation enumeration     applications/chem/algorithms/ucc.py:680-832, uccsd.py:268-318
* ``get_hop_from_integral``        chem_libs/hamiltonians_chem_library/hamiltonian_builders.py:71-108
* ``random_integral``              hamiltonian_builders.py:261-278
* reverse-sweep gradient (model)   chem_libs/quantum_chem_library/civector_ops.py:141-200

Third-party algorithm restated: OpenFermion 1.7.x ``jordan_wigner`` (a_k = prod_{j<k} Z_j (X_k+iY_k)/2)
and ``get_sparse_operator``; here orbital k <-> bit k of the flat index <-> engine qubit n-1-k
(SURVEY.md section 9).  The fermionic operators are applied directly to occupation-number
basis states, so this file is independent of the Pauli-sum code the CUDA host layer uses.

Parity pinning: ``evolve_excitation`` is checked against ``scipy.linalg.expm(theta*G)`` built from
the textbook JW matrices (tests/test_oracle_ucc.py), the H2 excitation list against the reference
docstring (uccsd.py:301-303).  Equality with PySCF's real H2O numbers: PARITY UNPINNED in this
environment (no pyscf/openfermion); the synthetic ``random_integral(7, 2077)`` problem has the same
shape (14 qubits, (5,5) electrons, 140 excitations / 75 parameters).
"""
from __future__ import annotations

from itertools import product
from typing import Callable, List, Sequence, Tuple

import numpy as np
import scipy.sparse as sp

from .sv_oracle import apply_kq, apply_1q, gate_x, init_statevector

C128 = np.complex128

# constants.py:5-15
_a = np.array([[0, 1], [0, 0]], dtype=np.float64)
_ad = _a.T
AD_A_HC = np.kron(_ad, _a) - np.kron(_ad, _a).T
_adad_aa = np.kron(np.kron(np.kron(_ad, _ad), _a), _a)
ADAD_AA_HC = _adad_aa - _adad_aa.T
AD_A_HC2 = AD_A_HC @ AD_A_HC
ADAD_AA_HC2 = ADAD_AA_HC @ ADAD_AA_HC


# --------------------------------------------------------------------------------------
# mini Jordan-Wigner algebra on Pauli-letter strings (only to derive Zset / sign)
# --------------------------------------------------------------------------------------
_MUL = {  # (a, b) -> (phase, a*b)
    ("I", "I"): (1, "I"), ("I", "X"): (1, "X"), ("I", "Y"): (1, "Y"), ("I", "Z"): (1, "Z"),
    ("X", "I"): (1, "X"), ("Y", "I"): (1, "Y"), ("Z", "I"): (1, "Z"),
    ("X", "X"): (1, "I"), ("Y", "Y"): (1, "I"), ("Z", "Z"): (1, "I"),
    ("X", "Y"): (1j, "Z"), ("Y", "X"): (-1j, "Z"),
    ("Y", "Z"): (1j, "X"), ("Z", "Y"): (-1j, "X"),
    ("Z", "X"): (1j, "Y"), ("X", "Z"): (-1j, "Y"),
}


def _pmul(A: dict, B: dict) -> dict:
    out: dict = {}
    for sa, ca in A.items():
        for sb, cb in B.items():
            ph = ca * cb
            letters = []
            for x, y in zip(sa, sb):
                p, l = _MUL[(x, y)]
                ph *= p
                letters.append(l)
            key = "".join(letters)
            out[key] = out.get(key, 0) + ph
    return {k: v for k, v in out.items() if abs(v) > 1e-14}


def _ladder(k: int, dagger: bool, n: int) -> dict:
    zs = "Z" * k
    tail = "I" * (n - k - 1)
    return {zs + "X" + tail: 0.5, zs + "Y" + tail: (-0.5j if dagger else 0.5j)}


def jw_excitation(f_idx: Sequence[int], n: int) -> dict:
    """JW image of a+_p a_q  or  a+_p a+_q a_r a_s (no h.c.), keys = letter strings over orbitals 0..n-1."""
    k = len(f_idx)
    ops = [_ladder(int(f_idx[i]), i < k // 2, n) for i in range(k)]
    acc = ops[0]
    for o in ops[1:]:
        acc = _pmul(acc, o)
    return acc


def excitation_zset_sign(f_idx: Sequence[int], n: int) -> Tuple[List[int], int]:
    """statevector_ops.py:45-54: Z positions of a JW term, and sign = +1 iff the coefficient of the
    lexicographically smallest term (sorted by ((idx, letter), ...) tuples) has positive real part."""
    qop = jw_excitation(f_idx, n)

    def key(s: str):
        return tuple((i, ch) for i, ch in enumerate(s) if ch != "I")

    first = min(qop.keys(), key=key)
    zset = [i for i, ch in enumerate(first) if ch == "Z"]
    for i, ch in enumerate(first):
        if ch not in "IZ":
            assert i in f_idx
    sign = 1 if qop[first].real > 0 else -1
    return zset, sign


# --------------------------------------------------------------------------------------
# excitation evolution
# --------------------------------------------------------------------------------------
def sign_vector(f_idx: Sequence[int], n: int) -> np.ndarray:
    """statevector_ops.py:55-62: sign * kron over tensor axes of [1,-1] on Z axes.  Orbital z sits
    on tensor axis n-1-z, i.e. on bit z of the flat index."""
    zset, sign = excitation_zset_sign(f_idx, n)
    i = np.arange(1 << n, dtype=np.int64)
    par = np.zeros(1 << n, dtype=np.int64)
    for z in zset:
        par ^= (i >> z) & 1
    return sign * (1.0 - 2.0 * par)


def apply_excitation(psi: np.ndarray, f_idx: Sequence[int], n: int, mode: str = "fermion") -> np.ndarray:
    """statevector_ops.py:25-63: G|psi> = (sign vector) * (local antisymmetric matrix on axes n-1-idx)."""
    qidx = [n - 1 - int(i) for i in f_idx]
    U = AD_A_HC if len(qidx) == 2 else ADAD_AA_HC
    out = apply_kq(psi, U, qidx, n)
    if mode != "fermion":
        return out
    return out * sign_vector(f_idx, n)


def evolve_excitation(psi: np.ndarray, f_idx: Sequence[int], theta: float, n: int, mode: str = "fermion") -> np.ndarray:
    """statevector_ops.py:140-168: psi + (1-cos t) G^2 psi + sin t G psi."""
    qidx = [n - 1 - int(i) for i in f_idx]
    U2 = AD_A_HC2 if len(qidx) == 2 else ADAD_AA_HC2
    f2 = apply_kq(psi, U2, qidx, n)
    f1 = apply_excitation(psi, f_idx, n, mode)
    return psi + (1.0 - np.cos(theta)) * f2 + np.sin(theta) * f1


def hf_state(n: int, n_elec_s: Tuple[int, int]) -> np.ndarray:
    """statevector_ops.py:172-199: x on wires n-1-i (i<nb) and n/2-1-i (i<na)."""
    na, nb = int(n_elec_s[0]), int(n_elec_s[1])
    psi = init_statevector(n)
    for i in range(nb):
        psi = apply_1q(psi, gate_x(), n - 1 - i, n)
    for i in range(na):
        psi = apply_1q(psi, gate_x(), n // 2 - 1 - i, n)
    return psi


def get_statevector(params: np.ndarray, n: int, n_elec_s, ex_ops, param_ids, mode: str = "fermion",
                    init_state: np.ndarray | None = None) -> np.ndarray:
    """statevector_ops.py:68-116."""
    psi = hf_state(n, n_elec_s) if init_state is None else np.asarray(init_state, dtype=C128)
    ids = param_ids if param_ids is not None else list(range(len(ex_ops)))
    for pid, f_idx in zip(ids, ex_ops):
        psi = evolve_excitation(psi, tuple(f_idx), float(params[pid]), n, mode)
    return psi


# --------------------------------------------------------------------------------------
# UCCSD excitation enumeration (ucc.py:680-832; uccsd.py:268-318 with init_method="zeros")
# --------------------------------------------------------------------------------------
def uccsd_ex_ops(no: int, nv: int) -> Tuple[List[tuple], List[int]]:
    """no/nv = occupied/virtual SPATIAL orbitals.  Spin-orbital numbering (ucc.py:768-778):
    beta-occ i, beta-virt no+a, alpha-occ no+nv+i, alpha-virt 2no+nv+a."""
    a_o = lambda i: no + nv + i
    a_v = lambda a: 2 * no + nv + a
    b_o = lambda i: i
    b_v = lambda a: no + a
    ex1: List[tuple] = []
    id1: List[int] = []
    pid = -1
    for i in range(no):
        for a in range(nv):
            pid += 1
            ex1 += [(a_v(a), a_o(i)), (b_v(a), b_o(i))]
            id1 += [pid, pid]
    ex2: List[tuple] = []
    id2: List[int] = []
    pid = -1
    for i in range(no):
        for j in range(i):
            for a in range(nv):
                for b in range(a):
                    pid += 1
                    ex2 += [(a_v(b), a_v(a), a_o(i), a_o(j)), (b_v(b), b_v(a), b_o(i), b_o(j))]
                    id2 += [pid, pid]
    for i in range(no):
        for j in range(i + 1):
            for a in range(nv):
                for b in range(a + 1):
                    if i == j and a == b:
                        pid += 1
                        ex2.append((b_v(a), a_v(a), a_o(i), b_o(i)))
                        id2.append(pid)
                        continue
                    pid += 1
                    ex2 += [(b_v(b), a_v(a), a_o(i), b_o(j)), (a_v(b), b_v(a), b_o(i), a_o(j))]
                    id2 += [pid, pid]
                    if i != j and a != b:
                        pid += 1
                        ex2 += [(b_v(a), a_v(b), a_o(i), b_o(j)), (a_v(a), b_v(b), b_o(i), a_o(j))]
                        id2 += [pid, pid]
    ex_ops = ex1 + ex2
    param_ids = id1 + [i + max(id1) + 1 for i in id2]
    return ex_ops, param_ids


# --------------------------------------------------------------------------------------
# Hamiltonian from integrals (hamiltonian_builders.py:71-108, 261-278)
# --------------------------------------------------------------------------------------
def random_integral(nao: int, seed: int = 2077) -> Tuple[np.ndarray, np.ndarray]:
    np.random.seed(seed)
    int1e = np.random.uniform(-1, 1, size=(nao, nao))
    int2e = np.random.uniform(-1, 1, size=(nao, nao, nao, nao))
    int1e = 0.5 * (int1e + int1e.T)
    int2e = 0.25 * (int2e + int2e.transpose((0, 1, 3, 2)) + int2e.transpose((1, 0, 2, 3))
                    + int2e.transpose((2, 3, 0, 1)))
    int2e = 0.5 * (int2e + int2e.transpose(3, 2, 1, 0))
    return int1e, int2e


def fermion_terms_from_integral(int1e: np.ndarray, int2e: np.ndarray) -> List[Tuple[Tuple[Tuple[int, int], ...], float]]:
    """List of (((orbital, dagger), ...), coeff) in operator order, as get_hop_from_integral builds them."""
    n_orb = int1e.shape[0]
    ns = 2 * n_orb
    h1e = np.zeros((ns, ns))
    h1e[:n_orb, :n_orb] = int1e
    h1e[n_orb:, n_orb:] = int1e
    h2e = np.zeros((ns, ns, ns, ns))
    for p, q, r, s in product(range(ns), repeat=4):
        if ((p < n_orb) == (s < n_orb)) and ((q < n_orb) == (r < n_orb)):
            h2e[p, q, r, s] = int2e[p % n_orb, s % n_orb, q % n_orb, r % n_orb]
    terms = []
    for p, q in product(range(ns), repeat=2):
        v = h1e[p, q]
        if abs(v) >= 1e-12:
            terms.append((((p, 1), (q, 0)), float(v)))
    for q, s in product(range(ns), repeat=2):
        for p, r in product(range(q), range(s)):
            v = h2e[p, q, r, s] - h2e[q, p, r, s]
            if abs(v) >= 1e-12:
                terms.append((((p, 1), (q, 1), (r, 0), (s, 0)), float(v)))
    return terms


def _apply_ladder(idx: np.ndarray, amp: np.ndarray, k: int, dagger: int) -> Tuple[np.ndarray, np.ndarray]:
    """a_k / a+_k on basis states |idx> (orbital k = bit k, JW sign = parity of occupied orbitals < k)."""
    occ = (idx >> k) & 1
    ok = (occ == 0) if dagger else (occ == 1)
    below = idx & ((1 << k) - 1)
    par = np.zeros_like(idx)
    b = below.copy()
    while np.any(b):
        par ^= b & 1
        b >>= 1
    sgn = 1.0 - 2.0 * par
    return idx ^ (1 << k), np.where(ok, amp * sgn, 0.0)


def fermion_sparse(terms, n: int) -> sp.csr_matrix:
    """sum_t c_t * (product of ladder operators) as a real 2^n x 2^n CSR matrix."""
    dim = 1 << n
    rows, cols, vals = [], [], []
    base = np.arange(dim, dtype=np.int64)
    for word, c in terms:
        idx = base.copy()
        amp = np.full(dim, float(c))
        for k, dg in reversed(word):  # rightmost operator acts first
            idx, amp = _apply_ladder(idx, amp, int(k), int(dg))
        nzm = amp != 0.0
        rows.append(idx[nzm]); cols.append(base[nzm]); vals.append(amp[nzm])
    H = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(dim, dim))
    return H.tocsr()


def hamiltonian_from_integral(int1e: np.ndarray, int2e: np.ndarray) -> sp.csr_matrix:
    return fermion_sparse(fermion_terms_from_integral(int1e, int2e), 2 * int1e.shape[0])


# --------------------------------------------------------------------------------------
# energy and gradient
# --------------------------------------------------------------------------------------
def energy(params, H, n, n_elec_s, ex_ops, param_ids, mode="fermion") -> float:
    """statevector_ops.py:224-234: Re sum conj(psi) * (H psi)."""
    psi = get_statevector(params, n, n_elec_s, ex_ops, param_ids, mode)
    return float(np.real(np.vdot(psi, H @ psi)))


def energy_and_grad_adjoint(params, H, n, n_elec_s, ex_ops, param_ids, mode="fermion") -> Tuple[float, np.ndarray]:
    """Analytic reverse sweep (civector_ops.py:141-200 carried over to the statevector):
    ket=psi_N, bra=H psi_N; for j=N..1: g[pid_j] += 2 Re<bra|G_j|ket>; ket,bra <- exp(-theta_j G_j)(ket,bra)."""
    params = np.asarray(params, dtype=np.float64)
    ket = get_statevector(params, n, n_elec_s, ex_ops, param_ids, mode)
    bra = H @ ket
    e = float(np.real(np.vdot(ket, bra)))
    g = np.zeros_like(params)
    for pid, f_idx in reversed(list(zip(param_ids, ex_ops))):
        g[pid] += 2.0 * float(np.real(np.vdot(bra, apply_excitation(ket, f_idx, n, mode))))
        ket = evolve_excitation(ket, f_idx, -float(params[pid]), n, mode)
        bra = evolve_excitation(bra, f_idx, -float(params[pid]), n, mode)
    return e, g


def energy_and_grad_forward_fd(params, H, n, n_elec_s, ex_ops, param_ids, mode="fermion", eps=1e-6):
    """The reference's numpy ``value_and_grad`` behaviour (numpy_backend.py:386-454): forward
    differences, P+1 energy evaluations."""
    params = np.asarray(params, dtype=np.float64)
    e0 = energy(params, H, n, n_elec_s, ex_ops, param_ids, mode)
    g = np.zeros_like(params)
    for i in range(params.size):
        p = params.copy(); p[i] += eps
        g[i] = (energy(p, H, n, n_elec_s, ex_ops, param_ids, mode) - e0) / eps
    return e0, g


# textbook JW matrices, used only to validate evolve_excitation against expm
def jw_generator_matrix(f_idx: Sequence[int], n: int) -> np.ndarray:
    """G = T - T^dagger with T = a+_p a_q or a+_p a+_q a_r a_s as a dense real 2^n x 2^n matrix."""
    k = len(f_idx)
    word = tuple((int(f_idx[i]), 1 if i < k // 2 else 0) for i in range(k))
    T = fermion_sparse([(word, 1.0)], n).toarray()
    return T - T.T
