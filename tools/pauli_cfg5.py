"""Config 5 pieces in isolation (run under gpurun): expectation values (tiled vs gather kernel) and sampling of 1024 x 20 qubits."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import PauliSum, _lib
from tyxonq_b200.batched import BatchedAnsatz

def med(fn, reps=5):
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), ts

dev = torch.device("cuda", 0); _lib.ensure_device(0)
nq, L, Bn, shots = 20, 4, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, 8192
params = np.random.default_rng(7).random((Bn, (L + 1) * nq))
terms = [(1.0, [(c, i), (c, i + 1)]) for i in range(nq - 1) for c in ("Z", "X", "Y")]
ham = PauliSum.from_pauli_list(nq, terms)
ba = BatchedAnsatz(nq, L, Bn, device=dev, dtype=torch.complex64)
ba.run(params); torch.cuda.synchronize()
a = ham.expectation(ba.state, tiled=True).real.cpu().numpy(); b = ham.expectation(ba.state, tiled=False).real.cpu().numpy()
print("max |tiled - gather| =", np.abs(a - b).max(), "mean E", a.mean())
t, ts = med(lambda: ham.expectation(ba.state, tiled=True)); print(f"tiled  {t:.2f} ms -> {Bn / t * 1e3:.0f} expvals/s, one-read fraction {Bn * (1 << nq) * 8 / t / 1e6 / 6557.8:.3f}", [f"{x:.1f}" for x in ts])
t, ts = med(lambda: ham.expectation(ba.state, tiled=False), 2); print(f"gather {t:.2f} ms -> {Bn / t * 1e3:.0f} expvals/s")
u = torch.from_numpy(np.random.default_rng(99).random((Bn, shots))).to(dev)
ba.sample(u); torch.cuda.synchronize()
t, ts = med(lambda: ba.sample(u)); print(f"sample {t:.2f} ms -> {Bn * shots / t * 1e3 / 1e6:.0f} M shots/s", [f"{x:.1f}" for x in ts])
