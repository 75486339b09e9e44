"""Energy + gradient evaluations per second of the small VQE configs (run under gpurun).
TFIM-10 (config 1): the CTA-resident kernel, one evaluation per call and batched; the fused-pass / CUDA-graph path beside it."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib
from tyxonq_b200.vqe import TFIMVqe

dev = torch.device("cuda", 0); _lib.ensure_device(0)
v = TFIMVqe(10, 1, device=dev)
p = np.random.default_rng(0).normal(size=(2, 10))
for name, kw in (("resident, one vector per call", {}), ("fused passes + CUDA graph (round 1)", {"resident": False})):
    for _ in range(3): v.energy_and_grad(p, **kw)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(200): e, g = v.energy_and_grad(p, **kw)
        ts.append((time.perf_counter() - t0) / 200)
    print(f"{name}: {1 / np.median(ts):.0f} evals/s ({1e6 * np.median(ts):.1f} us), E = {e:.12f}")
for B in (148, 1024, 8192):
    pb = np.random.default_rng(1).normal(size=(B, 20))
    v.energy_and_grad_batch(pb)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); eb, gb = v.energy_and_grad_batch(pb); ts.append(time.perf_counter() - t0)
    print(f"resident, batch {B}: {B / np.median(ts):.0f} evals/s ({1e3 * np.median(ts):.2f} ms per launch incl. H2D/D2H)")

# H2O-shaped UCCSD (config 2): one evaluation per call and concurrent replicas
from tyxonq_b200 import ucc
i1, i2 = ucc.random_integral(7, 2077)
ex_ops, pids = ucc.uccsd_ex_ops(5, 2)
sv = ucc.UCCStatevector(14, (5, 5), ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=dev)
np.random.seed(2077)
p = np.random.rand(75) - 0.5
for _ in range(3): sv.energy_and_grad(p)
ts = []
for _ in range(5):
    t0 = time.perf_counter()
    for _ in range(20): e, g = sv.energy_and_grad(p)
    ts.append((time.perf_counter() - t0) / 20)
print(f"H2O-shaped UCCSD, one evaluation per call: {1 / np.median(ts):.0f} evals/s, E = {e:.12f}")
pb = np.random.default_rng(5).uniform(-0.5, 0.5, (64, 75))
for R in (4, 8, 9):
    sv.energy_and_grad_batch(pb[:R], replicas=R)
    t0 = time.perf_counter(); es, gs = sv.energy_and_grad_batch(pb, replicas=R); dt = time.perf_counter() - t0
    print(f"H2O-shaped UCCSD, {R} concurrent replicas: {64 / dt:.0f} evals/s")
