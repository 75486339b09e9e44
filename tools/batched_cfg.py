"""Config 5 timing (run under gpurun): 20-qubit hwe-ry ansatz, L = 4, B parameter sets in one batched run."""
import sys, time
from pathlib import Path
import numpy as np, torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200.batched import BatchedAnsatz

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
ba = BatchedAnsatz(20, 4, B, device="cuda", dtype=torch.complex64)
params = np.random.default_rng(7).random((B, 100))
ba.run(params); torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    st = ba.run(params)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / 3
print(f"batched hwe20 x{B}: {B / dt:.0f} states/s, passes={ba.passes}, {dt * 1e3:.1f} ms, norm2[0]={float((st[0].abs() ** 2).sum()):.8f}")
