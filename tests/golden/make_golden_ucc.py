"""Generate tests/golden/reference_ucc.npz from the LIVE reference (run in the build container only).

The reference's chem stack (pyscf / openfermion) is not importable here, but everything exp(theta G) needs except the
Jordan-Wigner sign vector is: the local excitation matrices of applications/chem/constants.py:5-15 (loaded by file path:
the module imports only numpy) and the reference's own apply_kqubit_unitary (libs/quantum_library/kernels/
statevector.py:71-129).  evolve_excitation (chem_libs/quantum_chem_library/statevector_ops.py:140-168) in its
mode != "fermion" branch is exactly  psi + (1 - cos t) U2 psi + sin t U1 psi  with those two ingredients; this script
evaluates that formula with the reference's functions for single and double excitations on 8 qubits.

    PYTHONPATH=/root/reference/src python tests/golden/make_golden_ucc.py
"""
from __future__ import annotations

import importlib.util
from pathlib import Path

import numpy as np

REF = Path("/root/reference/src/tyxonq")
spec = importlib.util.spec_from_file_location("ref_chem_constants", REF / "applications" / "chem" / "constants.py")
const = importlib.util.module_from_spec(spec)
spec.loader.exec_module(const)

from tyxonq.libs.quantum_library.kernels.statevector import apply_kqubit_unitary  # noqa: E402
from tyxonq.numerics import get_backend  # noqa: E402

nb = get_backend("numpy")
n = 8
rng = np.random.default_rng(2026)
psi0 = rng.normal(size=1 << n) + 1j * rng.normal(size=1 << n)
psi0 /= np.linalg.norm(psi0)
ex_ops = [(4, 0), (6, 1), (5, 3), (7, 2), (6, 4, 0, 1), (7, 5, 2, 3), (7, 4, 1, 0), (5, 4, 3, 2)]
thetas = rng.uniform(-1.2, 1.2, len(ex_ops))


def evolve(psi, f_idx, theta):
    """statevector_ops.py:140-168 with mode != 'fermion' (statevector_ops.py:42-43): no sign vector."""
    qubit_idx = [n - 1 - int(i) for i in f_idx]
    U2, U1 = (const.ad_a_hc2, const.ad_a_hc) if len(qubit_idx) == 2 else (const.adad_aa_hc2, const.adad_aa_hc)
    f2 = apply_kqubit_unitary(psi, U2, qubit_idx, n, backend=nb)
    f1 = apply_kqubit_unitary(psi, U1, qubit_idx, n, backend=nb)
    return psi + (1.0 - np.cos(theta)) * f2 + np.sin(theta) * f1


singles = np.stack([evolve(psi0, f, t) for f, t in zip(ex_ops, thetas)])
seq = psi0
for f, t in zip(ex_ops, thetas):
    seq = evolve(seq, f, t)
out = Path(__file__).resolve().parent / "reference_ucc.npz"
np.savez(out, n=n, psi0=psi0, ex_ops=np.array([list(f) + [-1] * (4 - len(f)) for f in ex_ops]), thetas=thetas,
         each=singles, sequence=seq, ad_a_hc=const.ad_a_hc, adad_aa_hc=const.adad_aa_hc)
print("wrote", out, singles.shape, float(np.linalg.norm(seq)))
