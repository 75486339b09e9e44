"""Where the time of BatchedAnsatz.run (config 5: 1024 x 20 qubits, complex64) goes: host planning, upload, passes."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from tyxonq_b200 import _lib
from tyxonq_b200 import program as P
from tyxonq_b200.batched import BatchedAnsatz, hwe_ry_gates
from tyxonq_b200.fuse import fuse
from tyxonq_b200.planner import compile_program

nq, L, Bn = 20, 4, 1024
dev = torch.device("cuda:0")
params = np.random.default_rng(7).random((Bn, (L + 1) * nq))
ba = BatchedAnsatz(nq, L, Bn, device=dev, dtype=torch.complex64)
ba.run(params); torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); g = hwe_ry_gates(nq, L, params); t1 = time.perf_counter()
    gf = fuse(g); t2 = time.perf_counter()
    prog = compile_program(gf, nq, ba.tile, batch_mats=Bn, itemsize=8); t3 = time.perf_counter()
    dp = P.DeviceProgram(prog, dev, torch.complex64); torch.cuda.synchronize(); t4 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); dp.run(ba.state); e1.record(); torch.cuda.synchronize(); t5 = time.perf_counter()
    print(f"gates {1e3*(t1-t0):.1f} ms, fuse {1e3*(t2-t1):.1f}, plan {1e3*(t3-t2):.1f}, DeviceProgram (upload {dp.h2d_bytes/1e6:.1f} MB) {1e3*(t4-t3):.1f}, "
          f"run wall {1e3*(t5-t4):.1f} (device {e0.elapsed_time(e1):.1f} ms), passes {prog.n_passes}, jit {_lib.jit_stats()}")
t0 = time.perf_counter(); ba.run(params); torch.cuda.synchronize(); print("BatchedAnsatz.run", 1e3 * (time.perf_counter() - t0), "ms")
