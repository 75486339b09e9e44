"""The N > 1 path on CPU: world_size-2 and -4 gloo groups.  The planner (global<->local exchanges,
victim bit swaps, logical->physical map), the exchange itself and the rank-bit handling of diagonal
gates / controls are the product's; the local tile passes run through the host emulator of the CUDA
kernel (tests/emu) because this tier has no GPU.  Result must equal the oracle's full state."""
from __future__ import annotations

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sv_oracle as O


class _EmuLocal:
    """Oracle/emulator-backed local executor injected into ShardedState (test infrastructure)."""

    def init_local(self, state, n_local, global_base):
        state.zero_()
        if global_base == 0:
            state[0] = 1

    def run_local(self, state, gates, n_local, global_base, cache_slot):
        from tests.emu.emu import run_program_emulated
        from tyxonq_b200.planner import TileConfig, compile_program
        prog = compile_program(gates, n_local, TileConfig(m=min(n_local, 6), L=2), chain=True)
        out = run_program_emulated(prog, state.numpy(), global_base=global_base)
        state.copy_(torch.from_numpy(out))

    def zmasks_local(self, state, masks, global_base):
        p = np.abs(state.numpy()) ** 2
        idx = np.arange(p.size, dtype=np.int64) | int(global_base)
        return torch.tensor([float(np.sum(p * (1 - 2 * (np.array([bin(int(i) & m).count("1") & 1 for i in idx]))))) for m in masks], dtype=torch.float64)

    def reduce_local(self, state, n_local):
        p = np.abs(state.numpy()) ** 2
        idx = np.arange(p.size)
        z = np.array([np.sum(p * (1 - 2 * ((idx >> b) & 1))) for b in range(n_local)])
        return torch.from_numpy(z), torch.tensor(p.sum())

    def pauli_local(self, state, sub, global_base):
        """numpy statement of tqb_expect_pauli_sum on one shard: sum_j conj(psi_{j^x}) c (-1)^popc((base|j)&z) psi_j."""
        psi = state.numpy().astype(np.complex128)
        j = np.arange(psi.size, dtype=np.int64)
        tot = 0j
        for gi in range(sub.n_groups):
            x = int(sub.group_x[gi])
            assert x < psi.size, "xmask of a group handed to the local kernel must be local"
            for t in range(int(sub.group_ptr[gi]), int(sub.group_ptr[gi + 1])):
                zm = int(sub.term_z[t])
                par = np.array([bin((int(global_base) | int(i)) & zm).count("1") & 1 for i in j])
                tot += complex(sub.term_coef[t]) * np.sum(np.conj(psi[j ^ x]) * (1 - 2 * par) * psi)
        return torch.tensor([tot.real, tot.imag], dtype=torch.float64)


def _worker(rank, world, port, n, ops, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tyxonq_b200.sharded import ShardedState, lower_and_fuse, plan_sharded
        st = ShardedState(n, torch.complex128, torch.device("cpu"), backend=_EmuLocal())
        gates = lower_and_fuse(ops, n)
        plan = plan_sharded(gates, n, st.g)
        st.init_zero()
        st.run(plan)
        z = st.expect_z_all().numpy()
        zz = st.expect_zmasks([(1 << b) | (1 << (b + 1)) for b in range(n - 1)] + [(1 << n) - 1]).numpy()
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), st.state.numpy())
        if rank == 0:
            np.save(os.path.join(out_dir, "zz.npy"), zz)
        if rank == 0:
            np.save(os.path.join(out_dir, "z.npy"), z)
            np.save(os.path.join(out_dir, "phys.npy"), np.array(plan.final_phys))
            np.save(os.path.join(out_dir, "nex.npy"), np.array([plan.n_exchanges]))
    finally:
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,n,kind", [(2, 8, "hea"), (4, 9, "trotter"), (2, 7, "random"), (4, 8, "qaoa")])
def test_sharded_matches_oracle(tmp_path, world, n, kind):
    rng = np.random.default_rng(world * 10 + n)
    if kind == "hea":
        ops = O.hea_ops(n, 3, rng.uniform(-3, 3, 6 * n))
    elif kind == "trotter":
        ops = O.trotter_ops(*O.tfim_terms(n), 1.0, 2)
    elif kind == "qaoa":
        ops = O.qaoa_ring_ops(n, 2, rng.uniform(-3, 3, 4))
    else:
        from tests.conftest import random_ops
        ops = random_ops(rng, n, 80)
    mp.spawn(_worker, args=(world, _free_port(), n, ops, str(tmp_path)), nprocs=world, join=True)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    phys = np.load(tmp_path / "phys.npy")
    g = int(np.log2(world))
    n_local = n - g
    full_phys = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])  # physical index = rank bits | local
    # un-permute: logical index i sits at physical index with bit phys[l] = bit l of i
    idx = np.arange(1 << n)
    pidx = np.zeros(1 << n, dtype=np.int64)
    for l in range(n):
        pidx |= ((idx >> l) & 1) << int(phys[l])
    got = full_phys[pidx]
    assert np.abs(got - ref).max() < 1e-12
    z = np.load(tmp_path / "z.npy")
    for q in range(n):
        assert abs(z[n - 1 - q] - O.expect_z(ref, q, n)) < 1e-12
    zz = np.load(tmp_path / "zz.npy")
    p = np.abs(ref) ** 2
    for b in range(n - 1):
        sgn = 1 - 2 * (((idx >> b) & 1) ^ ((idx >> (b + 1)) & 1))
        assert abs(zz[b] - np.sum(p * sgn)) < 1e-12
    par = np.array([bin(int(i)).count("1") & 1 for i in idx])
    assert abs(zz[n - 1] - np.sum(p * (1 - 2 * par))) < 1e-12
    assert int(np.load(tmp_path / "nex.npy")[0]) >= 1  # the circuits touch the global qubits


def test_plan_needs_no_exchange_for_diagonal_or_control_use_of_global_bits():
    from tyxonq_b200.sharded import lower_and_fuse, plan_sharded
    n = 8
    ops = [("h", q) for q in range(2, n)] + [("rz", 0, 0.3), ("rzz", 0, 1, 0.2), ("cz", 1, 5), ("s", 0)]
    plan = plan_sharded(lower_and_fuse(ops, n), n, 2)
    assert plan.n_exchanges == 0
    ops2 = ops + [("h", 0)]
    assert plan_sharded(lower_and_fuse(ops2, n), n, 2).n_exchanges == 1


# ---- sharded sampling: restore the identity layout, all-gather the chunk totals, resolve locally ----
_BLK = 16  # chunk length of the test executor (the CUDA library uses 4096); the contract is the same for any length


class _EmuSampler(_EmuLocal):
    def chunk_totals(self, state, n_local):
        p = (state.numpy().real ** 2 + state.numpy().imag ** 2).reshape(-1, _BLK)
        return torch.from_numpy(np.cumsum(p, axis=1)[:, -1].copy())

    def chunk_prefix(self, totals_all):
        return torch.from_numpy(np.concatenate([[0.0], np.cumsum(totals_all.numpy())]))

    def sample_shard(self, state, n_local, prefix, chunk_first, tail_index, uniforms):
        pre = prefix.numpy()
        total = pre[-1]
        nc = pre.size - 1
        p = (state.numpy().real ** 2 + state.numpy().imag ** 2).reshape(-1, _BLK)
        out = np.empty(uniforms.numel(), dtype=np.int64)
        for s, u in enumerate(uniforms.numpy()):
            c = int(np.searchsorted(pre[1:] / total, u, side="right"))
            if c >= nc:
                out[s] = tail_index
            elif not (chunk_first <= c < chunk_first + p.shape[0]):
                out[s] = -1
            else:
                within = pre[c] + np.cumsum(p[c - chunk_first])
                out[s] = c * _BLK + int(np.searchsorted(within / total, u, side="right"))
        return torch.from_numpy(out)


def _sample_worker(rank, world, port, n, ops, uniforms, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tyxonq_b200.sharded import ShardedState, lower_and_fuse, plan_sharded
        st = ShardedState(n, torch.complex128, torch.device("cpu"), backend=_EmuSampler())
        plan = plan_sharded(lower_and_fuse(ops, n), n, st.g)
        st.init_zero()
        st.run(plan)
        moved = st.phys != list(range(n))
        idx = st.sample(torch.from_numpy(uniforms))
        assert st.phys == list(range(n))
        np.save(os.path.join(out_dir, f"idx{rank}.npy"), idx.numpy())
        np.save(os.path.join(out_dir, f"shard{rank}.npy"), st.state.numpy())
        counts = st.counts(torch.from_numpy(uniforms))   # a collective: every rank calls it
        if rank == 0:
            np.save(os.path.join(out_dir, "moved.npy"), np.array([int(moved)]))
            import json
            with open(os.path.join(out_dir, "counts.json"), "w") as f:
                json.dump(counts, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n,kind", [(2, 8, "hea"), (4, 9, "trotter"), (4, 8, "random")])
def test_sharded_sampling_is_bit_exact_with_the_oracle(tmp_path, world, n, kind):
    rng = np.random.default_rng(world * 100 + n)
    if kind == "hea":
        ops = O.hea_ops(n, 3, rng.uniform(-3, 3, 6 * n))
    elif kind == "trotter":
        ops = O.trotter_ops(*O.tfim_terms(n), 1.0, 2)
    else:
        from tests.conftest import random_ops
        ops = random_ops(rng, n, 80)
    u = rng.random(300)
    u[:3] = [0.0, 0.5, 1.0 - 2 ** -53]
    mp.spawn(_sample_worker, args=(world, _free_port(), n, ops, u, str(tmp_path)), nprocs=world, join=True)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    # after restore_layout the shards ARE the logical state, rank-major
    full = np.concatenate([np.load(tmp_path / f"shard{r}.npy") for r in range(world)])
    assert np.abs(full - ref).max() < 1e-12
    want = O.sample_indices(full.real ** 2 + full.imag ** 2, u, block=_BLK)
    for r in range(world):
        assert np.array_equal(np.load(tmp_path / f"idx{r}.npy"), want)
    assert int(np.load(tmp_path / "moved.npy")[0]) == 1   # the layout really had to be restored
    import json
    counts = json.load(open(tmp_path / "counts.json"))
    assert counts == O.counts_from_indices(want, n)


def test_plan_restore_reaches_identity_from_any_layout():
    from tyxonq_b200.sharded import plan_restore
    rng = np.random.default_rng(5)
    for n, g in [(8, 1), (9, 2), (10, 3), (7, 0)]:
        for _ in range(20):
            phys = list(rng.permutation(n))
            plan = plan_restore(phys, n, g)
            assert plan.final_phys == list(range(n)) and plan.n_exchanges <= 2
            # replay on an index table: apply bit swaps / exchanges to the physical positions
            cur = list(phys)
            n_local = n - g
            for seg in plan.segments:
                for gt in seg.gates:
                    a, b = gt.bits
                    assert a < n_local and b < n_local
                    cur = [b if p == a else a if p == b else p for p in cur]
                if seg.exchange_after:
                    m = {n_local - g + i: n_local + i for i in range(g)}
                    m.update({v: k for k, v in m.items()})
                    cur = [m.get(p, p) for p in cur]
            assert cur == list(range(n))


# ---- sharded Pauli sums: X / Y factors on rank bits need a layout change (or a basis rotation) first ----
def _pauli_worker(rank, world, port, n, ops, terms, weights, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tyxonq_b200.pauli import PauliSum
        from tyxonq_b200.sharded import ShardedState, lower_and_fuse, plan_sharded
        st = ShardedState(n, torch.complex128, torch.device("cpu"), backend=_EmuLocal())
        st.init_zero()
        st.run(plan_sharded(lower_and_fuse(ops, n), n, st.g))
        before = st.expect_z_all().numpy()
        vals, nex = [], []
        for tl, wl in zip(terms, weights):
            vals.append(complex(st.expect_pauli_sum(PauliSum.from_codes(tl, wl))))
            nex.append(st.pauli_exchanges)
        after = st.expect_z_all().numpy()
        assert np.abs(before - after).max() < 1e-12     # the state is the same state afterwards
        # and gates can still be applied in whatever layout it is left in
        st.apply(lower_and_fuse([("h", 0), ("cx", 0, n - 1)], n))
        z2 = st.expect_z_all().numpy()
        if rank == 0:
            np.save(os.path.join(out_dir, "vals.npy"), np.array(vals))
            np.save(os.path.join(out_dir, "nex.npy"), np.array(nex))
            np.save(os.path.join(out_dir, "z2.npy"), z2)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 7), (4, 8)])
def test_sharded_pauli_sum_matches_oracle(tmp_path, world, n):
    rng = np.random.default_rng(100 * world + n)
    ops = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n))
    heis_t, heis_w = O.heisenberg_terms(n, [(i, (i + 1) % n) for i in range(n)])
    rand_t = [list(rng.integers(0, 4, n)) for _ in range(12)]
    rand_w = list(rng.uniform(-1, 1, 12))
    wide_t = [[1] * n, [2] * n, [1, 2] * (n // 2) + [1] * (n % 2), [2] + [1] * (n - 1), [3] * n, [0] * n]
    wide_w = [0.7, -0.4, 0.3, 0.2, 0.5, 1.5]
    diag_t, diag_w = [[3 if q in (0, 1) else 0 for q in range(n)], [3] + [0] * (n - 1)], [0.9, -0.2]
    sets = [(heis_t, list(heis_w)), (rand_t, rand_w), (wide_t, wide_w), (diag_t, diag_w)]
    mp.spawn(_pauli_worker, args=(world, _free_port(), n, ops, [s[0] for s in sets], [s[1] for s in sets], str(tmp_path)),
             nprocs=world, join=True)
    ref, _ = O.evolve_ops(n, ops, mode="run")
    vals = np.load(tmp_path / "vals.npy")
    for (tl, wl), v in zip(sets, vals):
        assert abs(v.imag) < 1e-12
        assert abs(v.real - O.expect_pauli_sum(ref, tl, wl)) < 1e-12
    nex = np.load(tmp_path / "nex.npy")
    assert nex[3] == 0                                   # diagonal Hamiltonian: no exchange whatever the layout
    ref2, _ = O.evolve_ops(n, list(ops) + [("h", 0), ("cx", 0, n - 1)], mode="run")
    z2 = np.load(tmp_path / "z2.npy")
    for q in range(n):
        assert abs(z2[n - 1 - q] - O.expect_z(ref2, q, n)) < 1e-12


def test_plan_localize_one_exchange_and_victim_choice():
    from tyxonq_b200.sharded import plan_localize
    n, g = 8, 2
    phys = list(range(n))
    assert plan_localize(phys, 0b00111111, n, g).segments == []          # already local
    cost = [5, 0, 5, 5, 0, 5, 9, 9]
    plan = plan_localize(phys, 0b11000000, n, g, cost)
    assert plan.n_exchanges == 1 and len(plan.segments) == 1
    n_local = n - g
    assert all(plan.final_phys[b] < n_local for b in (6, 7))
    assert sorted(l for l in range(n) if plan.final_phys[l] >= n_local) == [1, 4]   # the cheapest bits went away
    with pytest.raises(RuntimeError):
        plan_localize(phys, 0b11111110, n, g)
    # six flipped bits, one of them on a rank bit, and only ONE spare local bit: the other spare bit is itself on a
    # rank bit, so it must come home first (two exchanges)
    mask = 0b10111110
    plan = plan_localize(phys, mask, n, g)
    assert plan.n_exchanges == 2
    assert all(plan.final_phys[b] < n_local for b in range(n) if (mask >> b) & 1)
    rng = np.random.default_rng(3)
    for _ in range(300):
        n2, g2 = int(rng.integers(8, 13)), int(rng.integers(1, 4))
        ph = [int(v) for v in rng.permutation(n2)]
        bits = rng.choice(n2, size=int(rng.integers(0, n2 - g2 + 1)), replace=False)
        mk = sum(1 << int(b) for b in bits)
        pl = plan_localize(ph, mk, n2, g2)
        assert sorted(pl.final_phys) == list(range(n2)) and all(pl.final_phys[int(b)] < n2 - g2 for b in bits)


# ---- seam B1 over ranks: ShardedStatevectorEngine.run / expval, called collectively ----
def _engine_worker(rank, world, port, n, ops, uniforms, ham, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import json
        from tests.conftest import FakeCircuit
        from tyxonq_b200 import ShardedStatevectorEngine
        eng = ShardedStatevectorEngine(device="cpu", local_backend=_EmuSampler())
        c = FakeCircuit(n, ops)
        r_counts = eng.run(c, shots=len(uniforms), uniforms=uniforms)
        r_exp = eng.run(c, shots=0)
        r_seed = eng.run(c, shots=64, seed=11)          # rank 0 draws, everyone gets the same uniforms
        e = eng.expval(c, ham)
        amps = [eng.amplitude(c, format(i, f"0{n}b")) for i in (0, 1, (1 << n) - 1, 0b1011 << (n - 4), 37)]
        with pytest.raises(NotImplementedError):
            eng.run(FakeCircuit(n, list(ops) + [("reset", 0)]), shots=0)
        with pytest.raises(NotImplementedError):
            eng.state(c)                                 # never gathered; the driver's shots == 0 epilogue gets no statevector
        shard, base = eng.state_shard(c)
        assert base == rank << (n - eng.last_state.g)
        np.save(os.path.join(out_dir, f"eshard{rank}.npy"), shard.numpy())
        with open(os.path.join(out_dir, f"res{rank}.json"), "w") as f:
            json.dump({"counts": r_counts, "exp": r_exp, "seed": r_seed, "e": e, "nex": eng.last_exchanges,
                       "amps": [[a.real, a.imag] for a in amps]}, f)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 8), (4, 9)])
def test_sharded_engine_run_contract(tmp_path, world, n):
    import json
    rng = np.random.default_rng(7 * world + n)
    ops = O.hea_ops(n, 2, rng.uniform(-3, 3, 4 * n)) + [("cry", 0, n - 1, 0.7), ("measure_z", 0), ("rxx", 1, n - 2, 0.4)]
    ops += [("measure_z", q) for q in range(1, n)]
    u = rng.random(200)
    ham = [(0.5, [("X", 0), ("X", 1)]), (-0.3, [("Y", 0), ("Z", n - 1)]), (0.8, [("Z", 1)]), (0.25, []), (0.4, [("Y", n - 1), ("Y", n - 2)])]
    mp.spawn(_engine_worker, args=(world, _free_port(), n, ops, u, ham, str(tmp_path)), nprocs=world, join=True)
    res = [json.load(open(tmp_path / f"res{r}.json")) for r in range(world)]
    assert all(r == res[0] for r in res[1:])            # every rank holds the same result
    ref, measures = O.evolve_ops(n, ops, mode="run")
    want = O.sample_indices(ref.real ** 2 + ref.imag ** 2, u, block=_BLK)
    assert res[0]["counts"]["result"] == O.counts_from_indices(want, n)
    assert res[0]["counts"]["metadata"] == {"shots": 200, "backend": "b200-sharded", "three_level": False}
    exp = O.run_expectations(n, ops)
    assert set(res[0]["exp"]["expectations"]) == set(exp)
    for k, v in exp.items():
        assert abs(res[0]["exp"]["expectations"][k] - v) < 1e-12
    want64 = O.sample_indices(ref.real ** 2 + ref.imag ** 2, np.random.default_rng(11).random(64), block=_BLK)
    assert res[0]["seed"]["result"] == O.counts_from_indices(want64, n)
    codes = {"X": 1, "Y": 2, "Z": 3}
    terms, w = [], []
    for c, lst in ham:
        t = [0] * n
        for p_, q in lst:
            t[q] = codes[p_]
        terms.append(t); w.append(c)
    # expval follows engine.state(): the op set of "state" mode (no cry, engine.py:918-1038)
    psi_state, _ = O.evolve_ops(n, ops, mode="state")
    assert abs(res[0]["e"] - O.expect_pauli_sum(psi_state, terms, w)) < 1e-12
    assert res[0]["nex"] >= 1
    full = np.concatenate([np.load(tmp_path / f"eshard{r}.npy") for r in range(world)])
    assert np.abs(full - psi_state).max() < 1e-12
    for i, (re, im) in zip((0, 1, (1 << n) - 1, 0b1011 << (n - 4), 37), res[0]["amps"]):
        assert abs(complex(re, im) - psi_state[i]) < 1e-12


def _grad_worker(rank, world, port, n, template, ham, params, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tyxonq_b200 import ShardedStatevectorEngine
        eng = ShardedStatevectorEngine(device="cpu", local_backend=_EmuLocal())
        e, g = eng.energy_and_grad(n, template, ham, params)
        np.save(os.path.join(out_dir, f"eg{rank}.npy"), np.concatenate([[e], g]))
    finally:
        dist.destroy_process_group()


def test_sharded_parameter_shift_gradient(tmp_path):
    from tyxonq_b200.vqe import Param
    n, world = 6, 2
    template = [("h", q) for q in range(n)] + [("ry", 0, Param(0)), ("rx", 1, Param(1)), ("cx", 0, 1), ("rz", 2, Param(2)),
                ("rzz", 0, 5, Param(3)), ("cx", 2, 3), ("ry", 3, Param(1, 2.0)), ("rxx", 0, 4, Param(0, -1.0)), ("cx", 4, 5), ("rx", 5, Param(2))]
    ham = [(0.5, [("X", 0), ("X", 1)]), (-0.3, [("Y", 0), ("Z", 5)]), (0.8, [("Z", 1)]), (0.25, []), (0.4, [("Y", 4), ("Y", 3)]), (0.6, [("Z", 0), ("Z", 2)])]
    params = np.random.default_rng(9).uniform(-1, 1, 4)
    mp.spawn(_grad_worker, args=(world, _free_port(), n, template, ham, params, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"eg{r}.npy") for r in range(world)]
    assert np.array_equal(got[0], got[1])
    codes = {"X": 1, "Y": 2, "Z": 3}
    terms, w = [], []
    for c, lst in ham:
        t = [0] * n
        for p_, q in lst:
            t[q] = codes[p_]
        terms.append(t); w.append(c)

    def energy(th):
        ops = [tuple(a.scale * float(th[a.index]) if isinstance(a, Param) else a for a in op) for op in template]
        psi, _ = O.evolve_ops(n, ops, mode="state")
        return O.expect_pauli_sum(psi, terms, w)

    assert abs(got[0][0] - energy(params)) < 1e-12
    assert np.abs(got[0][1:] - O.central_fd_gradient(energy, params, eps=1e-5)).max() < 1e-8
