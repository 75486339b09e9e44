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
    dist.destroy_process_group()
    if rank == 0 and not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
