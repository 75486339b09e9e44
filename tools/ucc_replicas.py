"""H2O-shaped UCCSD energy + gradient: throughput against the number of concurrent replicas (run under gpurun)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import ucc

dev = torch.device("cuda", 0)
i1, i2 = ucc.random_integral(7, 2077)
ex_ops, pids = ucc.uccsd_ex_ops(5, 2)
sv = ucc.UCCStatevector(14, (5, 5), ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=dev)
p = np.random.default_rng(5).uniform(-0.5, 0.5, (96, 75))
ref = None
for R in (1, 8, 16, 24, 32):
    sv.energy_and_grad_batch(p[:R], replicas=R)
    t0 = time.perf_counter()
    es, gs = sv.energy_and_grad_batch(p, replicas=R)
    dt = time.perf_counter() - t0
    if ref is None:
        ref = (es.copy(), gs.copy())
    print(f"replicas {R:2d}: {96 / dt:8.0f} evaluations/s, max |dE| {np.abs(es - ref[0]).max():.1e}, max |dg| {np.abs(gs - ref[1]).max():.1e}", flush=True)
