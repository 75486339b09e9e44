"""Energy + gradient of the hardware-efficient ansatz at scale (run under gpurun): the layered adjoint sweep against a forward
evaluation.  usage: grad_scale.py n layers"""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib
from tyxonq_b200.pauli import PauliSum
from tyxonq_b200.vqe import AdjointEnergy, Param

n, layers = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda", 0); _lib.ensure_device(0)
ops = [("h", q) for q in range(n)]; k = 0
for _ in range(layers):
    ops += [("cx", q, q + 1) for q in range(n - 1)]
    for q in range(n):
        ops.append(("rz", q, Param(k))); k += 1
        ops.append(("rx", q, Param(k))); k += 1
ham = PauliSum.from_pauli_list(n, [(1.0, [("Z", i), ("Z", i + 1)]) for i in range(n - 1)] + [(-1.0, [("X", i)]) for i in range(n)])
ae = AdjointEnergy(n, ops, ham, device=dev)
th = np.random.default_rng(0).uniform(-np.pi, np.pi, k)
_lib.load().tqb_set_jit(1)
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e = ae.energy(th)
    torch.cuda.synchronize(); t_f = time.perf_counter() - t0
    t0 = time.perf_counter()
    e2, g = ae.energy_and_grad_layered(th)
    torch.cuda.synchronize(); t_g = time.perf_counter() - t0
    _lib.load().tqb_jit_wait()
    print(f"n={n} layers={layers} params={k}: energy {t_f:.2f} s, energy+gradient (layered) {t_g:.2f} s = {t_g / t_f:.1f} x, |E diff| {abs(e - e2):.1e}, |g|max {np.abs(g).max():.3f}", flush=True)
