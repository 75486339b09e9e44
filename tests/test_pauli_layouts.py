"""Host logic of the tile-staged Pauli sums: every xmask group lands in exactly one layout whose tile bits contain it."""
from __future__ import annotations

import numpy as np

from tyxonq_b200.pauli import PauliSum


def test_layouts_cover_every_group():
    rng = np.random.default_rng(4)
    for n, m, l_min in ((20, 13, 4), (14, 12, 3), (30, 12, 3), (8, 8, 3)):
        xs = [0] + [int(sum(1 << int(b) for b in rng.choice(n, size=int(rng.integers(1, 5)), replace=False))) for _ in range(50)]
        lay = PauliSum.plan_layouts(xs, n, m, l_min)
        assert lay is not None
        seen = sorted(g for _, gs in lay for g in gs)
        assert seen == list(range(len(xs)))
        for bits, gs in lay:
            assert len(bits) == min(m, n) and bits == sorted(bits) and bits[:min(l_min, m)] == list(range(min(l_min, m)))
            mask = sum(1 << b for b in bits)
            assert all(xs[g] & ~mask == 0 for g in gs)


def test_heisenberg_chain_needs_two_layouts():
    """Config 5: 57 terms on 20 qubits, nearest-neighbour XX / YY / ZZ: two reads of the state with 64 KiB complex64 tiles."""
    n = 20
    terms = [(1.0, [(c, i), (c, i + 1)]) for i in range(n - 1) for c in ("Z", "X", "Y")]
    ham = PauliSum.from_pauli_list(n, terms)
    assert ham.hermitian and ham.real_coef and ham.n_groups == 20
    lay = PauliSum.plan_layouts([int(x) for x in ham.group_x], n, 13, 4)
    assert lay is not None and len(lay) == 2


def test_unfittable_xmask_returns_none():
    assert PauliSum.plan_layouts([(1 << 20) - 1], 20, 12, 3) is None
