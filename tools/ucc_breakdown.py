"""H2O-shaped UCCSD evaluation: device time of its parts (forward sweep, H|psi>, <psi|H|psi>, reverse sweep)."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib, ucc
from tyxonq_b200 import program as P

dev = torch.device("cuda", 0)
i1, i2 = ucc.random_integral(7, 2077)
ex_ops, pids = ucc.uccsd_ex_ops(5, 2)
sv = ucc.UCCStatevector(14, (5, 5), ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=dev)
p = np.random.default_rng(5).uniform(-0.5, 0.5, 75)
sv.energy_and_grad(p); sv.energy_and_grad(p)
lib = _lib.load()
kb = sv._kb
ptr, n, _, dt, stream = P._prep(kb[0])
N = len(sv._proto)
step_bytes = sv._steps_np.dtype.itemsize

def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

print("terms", sv.ham.n_terms, "groups", sv.ham.n_groups, "excitations", N, "use_sweep", sv.use_sweep)
print(f"forward sweep   {t(lambda: lib.tqb_pair_sweep(kb[0].data_ptr(), 0, n, dt, sv._steps_dev.data_ptr(), N, 0, 0, sv._sync.data_ptr(), stream)):8.1f} us")
print(f"H|psi>          {t(lambda: sv.ham.apply(kb[0], kb[1])):8.1f} us")
print(f"<ket|bra>       {t(lambda: lib.tqb_inner(kb[0].data_ptr(), kb[1].data_ptr(), n, 1, dt, sv._e.data_ptr(), stream)):8.1f} us")
print(f"reverse sweep   {t(lambda: lib.tqb_pair_sweep(kb[0].data_ptr(), kb[1].data_ptr(), n, dt, sv._steps_dev.data_ptr() + N * step_bytes, N, 1, sv._gout.data_ptr(), sv._sync.data_ptr() + 8, stream)):8.1f} us")
t0 = time.perf_counter()
for _ in range(50): sv.energy_and_grad(p)
print(f"energy_and_grad {(time.perf_counter() - t0) / 50 * 1e6:8.1f} us per call (graph replay + host)")
