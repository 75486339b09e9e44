"""Multi-GPU correctness check (torchrun, one rank per GPU): the sharded run of a circuit must reproduce the
single-GPU state of the same circuit (checked through <Z_q> for every qubit, the norm, and sampled amplitudes).
usage: torchrun --nproc-per-node N tools/sharded_check.py [n] [layers]"""
from __future__ import annotations

import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.circuits import Circuit, hea_ops, trotter_ops, tfim_terms  # noqa: E402
from tyxonq_b200.engine import StatevectorEngine  # noqa: E402
from tyxonq_b200.pauli import PauliSum  # noqa: E402
from tyxonq_b200.sharded import ShardedState, lower_and_fuse, plan_sharded  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 26
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    _lib.ensure_device(lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for name, ops in (("hea", hea_ops(n, layers, np.random.default_rng(1).uniform(-3, 3, 2 * layers * n))),
                      ("trotter", [o for o in trotter_ops(*tfim_terms(n), 1.0, 2) if o[0] != "measure_z"])):
        for dt, tol in ((torch.complex128, 1e-10), (torch.complex64, 1e-4)):
            st = ShardedState(n, dt, dev)
            plan = plan_sharded(lower_and_fuse(ops, n), n, st.g)
            st.init_zero()
            st.run(plan)
            z = st.expect_z_all().cpu().numpy()
            nrm = P.norm2(st.state)
            dist.all_reduce(nrm)
            # sharded sampling: restore the identity layout, then the blocked-CDF sampler across ranks.  It must return
            # exactly what the single-GPU sampler returns on the gathered state (same contract, bit for bit).
            u = torch.from_numpy(np.random.default_rng(3).random(4096))
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            idx = st.sample(u)
            t1.record(); torch.cuda.synchronize()
            sample_ms = t0.elapsed_time(t1)
            parts = [torch.empty_like(st.state) for _ in range(world)] if rank == 0 else None
            dist.gather(torch.view_as_real(st.state), [torch.view_as_real(x) for x in parts] if rank == 0 else None, dst=0)
            # sharded Pauli sum with X / Y factors on every qubit (rank bits included): ring Heisenberg + fields
            ham = PauliSum.from_pauli_list(n, [(c, [(p_, q), (p_, (q + 1) % n)]) for q in range(n) for p_, c in (("X", 0.5), ("Y", -0.3), ("Z", 0.8))]
                                           + [(0.2, [("X", q)]) for q in range(n)] + [(0.1, [("Y", 0), ("Z", n // 2), ("X", n - 1)])])
            e_sh = complex(st.expect_pauli_sum(ham))
            n_pex = st.pauli_exchanges
            if rank == 0:
                full = torch.cat(parts)
                del parts
                same = bool(torch.equal(P.sample(full, u), idx))
                eng = StatevectorEngine("b200", device=dev, dtype=dt)
                ref, _, _ = eng._evolve(Circuit(n, ops), "run")
                zr = P.expect_z_bits(ref)[0].cpu().numpy()
                err = float(np.abs(z - zr).max())
                amp_err = float((full - ref).abs().max())
                e_ref = complex(ham.expectation(ref).cpu().numpy()[0])
                e_err = abs(e_sh - e_ref)
                good = err < tol and abs(float(nrm[0]) - 1.0) < tol and same and amp_err < tol and e_err < tol * n
                ok &= good
                print(f"{name} n={n} world={world} {dt}: exchanges={plan.n_exchanges} segments={len(plan.segments)} max|dZ|={err:.2e} "
                      f"norm-1={float(nrm[0]) - 1.0:.2e} restored-state max|d|={amp_err:.2e} sharded sampler == single-GPU sampler: {same} "
                      f"(restore + 4096 shots {sample_ms:.1f} ms) sharded <H> err={e_err:.2e} ({n_pex} exchanges) {'OK' if good else 'FAIL'}", flush=True)
                del ref, full
            del st
            torch.cuda.empty_cache()
            dist.barrier()
    # seam B1 over ranks: ShardedStatevectorEngine.run / expval against the single-GPU engine on the same circuit
    from tyxonq_b200 import ShardedStatevectorEngine
    ops = hea_ops(n, layers, np.random.default_rng(2).uniform(-3, 3, 2 * layers * n)) + [("measure_z", q) for q in range(n)]
    circ = Circuit(n, ops)
    u = np.random.default_rng(5).random(2048)
    ham = [(0.5, [("X", q), ("X", (q + 1) % n)]) for q in range(n)] + [(-0.3, [("Y", 0), ("Z", n - 1)]), (0.25, [])] \
        + [(0.4, [("Z", q)]) for q in range(n)]
    seng = ShardedStatevectorEngine(device=dev)
    r_counts = seng.run(circ, shots=2048, uniforms=u)
    r_exp = seng.run(circ, shots=0)
    e_sh = seng.expval(circ, ham)
    if rank == 0:
        eng = StatevectorEngine("b200", device=dev)
        w_counts = eng.run(circ, shots=2048, uniforms=u)
        w_exp = eng.run(circ, shots=0)
        e_one = eng.expval(circ, ham)
        same_counts = r_counts["result"] == w_counts["result"]
        dz = max(abs(r_exp["expectations"][k] - v) for k, v in w_exp["expectations"].items())
        good = same_counts and dz < 1e-10 and abs(e_sh - e_one) < 1e-10 and set(r_exp["expectations"]) == set(w_exp["expectations"])
        ok &= good
        print(f"ShardedStatevectorEngine n={n} world={world}: counts == single-GPU counts: {same_counts} ({len(r_counts['result'])} bitstrings), "
              f"max|dZ|={dz:.2e}, expval err={abs(e_sh - e_one):.2e}, exchanges={seng.last_exchanges} {'OK' if good else 'FAIL'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
