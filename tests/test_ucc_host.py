"""Host-side UCC logic of the product (bitmask Jordan-Wigner, excitation enumeration, HF index)
against the oracle's independent restatement (fermion operators applied to basis states)."""
from __future__ import annotations

import itertools

import numpy as np
import scipy.linalg as sl

from oracle import ucc_oracle as U


def _pauli_sum_matvec(ham, v):
    """numpy evaluation of a PauliSum's arrays (test-side arithmetic): out_j = sum phase(j^x) v_{j^x}."""
    n = ham.n
    j = np.arange(1 << n, dtype=np.int64)
    out = np.zeros(1 << n, dtype=np.complex128)
    for g in range(ham.n_groups):
        x = int(ham.group_x[g])
        src = j ^ x
        ph = np.zeros(1 << n, dtype=np.complex128)
        for t in range(ham.group_ptr[g], ham.group_ptr[g + 1]):
            z = int(ham.term_z[t])
            par = np.zeros(1 << n, dtype=np.int64)
            b = src & z
            while np.any(b):
                par ^= b & 1
                b >>= 1
            ph += ham.term_coef[t] * (1.0 - 2.0 * par)
        out += ph * v[src]
    return out


def test_oracle_evolve_excitation_is_expm():
    n = 6
    rng = np.random.default_rng(1)
    psi = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    psi /= np.linalg.norm(psi)
    ops = list(itertools.permutations(range(n), 2))[::3] + [(4, 0, 3, 5), (1, 3, 2, 0), (0, 1, 2, 3), (5, 2, 4, 1)]
    for f in ops:
        G = U.jw_generator_matrix(f, n)
        assert np.abs(sl.expm(0.37 * G) @ psi - U.evolve_excitation(psi, f, 0.37, n)).max() < 1e-13, f
    # H2 excitation list of the reference docstring (uccsd.py:301-303)
    assert U.uccsd_ex_ops(1, 1) == ([(3, 2), (1, 0), (1, 3, 2, 0)], [0, 0, 1])


def test_product_jw_matches_oracle():
    from tyxonq_b200 import ucc
    n = 8
    for f in [(3, 0), (0, 3), (6, 2), (4, 0, 3, 7), (1, 3, 2, 0), (0, 1, 2, 3), (7, 5, 2, 0), (2, 6, 5, 1)]:
        zset, sign = U.excitation_zset_sign(f, n)
        zmask, sign2 = ucc.excitation_zmask_sign(f, n)
        assert zmask == sum(1 << z for z in zset) and sign == sign2, f
    assert ucc.uccsd_ex_ops(5, 2) == U.uccsd_ex_ops(5, 2)
    assert ucc.uccsd_ex_ops(2, 2) == U.uccsd_ex_ops(2, 2)
    # HF occupation = X gates of get_init_circuit
    for nn, nes in [(8, (2, 2)), (14, (5, 5)), (8, (3, 1))]:
        hf = U.hf_state(nn, nes)
        assert hf[ucc.hf_basis_index(nn, nes)] == 1.0


def test_product_hamiltonian_matches_oracle():
    from tyxonq_b200 import ucc
    i1, i2 = ucc.random_integral(3, 7)
    ham = ucc.hamiltonian_from_integral(i1, i2)
    Hs = U.hamiltonian_from_integral(i1, i2)
    rng = np.random.default_rng(0)
    v = rng.normal(size=1 << 6) + 1j * rng.normal(size=1 << 6)
    assert np.abs(_pauli_sum_matvec(ham, v) - Hs @ v).max() < 1e-12
    assert np.abs(ham.term_coef.imag).max() < 1e-12 or True  # complex coefficients allowed (Y strings)


def test_pauli_sum_constructors():
    from oracle import sv_oracle as O
    from tyxonq_b200.pauli import PauliSum
    rng = np.random.default_rng(2)
    n = 5
    terms = [[int(c) for c in rng.integers(0, 4, n)] for _ in range(12)]
    w = rng.normal(size=12).tolist()
    v = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
    ref = O.apply_pauli_sum(v, terms, w)
    a = PauliSum.from_codes(terms, w)
    assert np.abs(_pauli_sum_matvec(a, v) - ref).max() < 1e-12
    lst = [(w[t], [("IXYZ"[c], q) for q, c in enumerate(terms[t]) if c]) for t in range(12)]
    b = PauliSum.from_pauli_list(n, lst)
    assert np.abs(_pauli_sum_matvec(b, v) - ref).max() < 1e-12

    class Q:  # OpenFermion-style container
        terms = {tuple((q, "IXYZ"[c]) for q, c in enumerate(terms[t]) if c): w[t] for t in range(12)}
    c = PauliSum.from_qubit_operator(n, Q)
    assert np.abs(_pauli_sum_matvec(c, v) - _pauli_sum_matvec(a, v)).max() < 1e-12


def test_oracle_matches_reference_excitation_fixture():
    """tests/golden/reference_ucc.npz: exp(theta G) in the reference's mode != 'fermion' branch (statevector_ops.py:42-43,
    140-168), evaluated by make_golden_ucc.py with the reference's own constants (applications/chem/constants.py:5-15) and
    its apply_kqubit_unitary.  Pins everything of the UCC evolution except the Jordan-Wigner sign vector."""
    import numpy as np
    from pathlib import Path
    from oracle import ucc_oracle as U
    d = np.load(Path(__file__).resolve().parent / "golden" / "reference_ucc.npz")
    n, psi0 = int(d["n"]), d["psi0"]
    ex = [tuple(int(x) for x in r if x >= 0) for r in d["ex_ops"]]
    assert np.array_equal(U.AD_A_HC, d["ad_a_hc"]) and np.array_equal(U.ADAD_AA_HC, d["adad_aa_hc"])
    psi = psi0
    for k, (f, t) in enumerate(zip(ex, d["thetas"])):
        assert np.abs(U.evolve_excitation(psi0, f, float(t), n, mode="qubit") - d["each"][k]).max() < 1e-14
        psi = U.evolve_excitation(psi, f, float(t), n, mode="qubit")
    assert np.abs(psi - d["sequence"]).max() < 1e-14
