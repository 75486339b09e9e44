"""Incremental cost of one in-tile gate sweep (run under gpurun): a single pass carrying G gates of one
kind on tile-local bits, G = 1, 2, 4, 8, 16.  usage: gate_cost.py n dtype m L threads kind(dense|mux|swap|diag|dense2)"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200 import gates as G  # noqa: E402
from tyxonq_b200.planner import TileConfig, compile_program  # noqa: E402


def main():
    n, dt, m, L, thr, kind = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), sys.argv[6]
    tdt = torch.complex128 if dt == "c128" else torch.complex64
    B = 16 if dt == "c128" else 8
    dev = torch.device("cuda", 0)
    _lib.ensure_device(0)
    st = P.new_state(n, dtype=tdt, device=dev)
    rng = np.random.default_rng(0)
    res = []
    for cnt in (1, 2, 4, 8, 16):
        lg = []
        for i in range(cnt):
            b = i % L  # low, always tile-local bit
            b2 = (i + 1) % L
            u = G.rx_mat(float(rng.uniform(-3, 3)))
            if kind == "dense":
                lg.append(G.LGate(G.DENSE, (b,), u))
            elif kind == "mux":
                lg.append(G.mux_gate(u, G.X_MAT @ u, b, b2))
            elif kind == "swap":
                lg.append(G.LGate(G.SWAP, (b, b2), G.X_MAT.reshape(4).copy(), pat_a=2, pat_b=3))
            elif kind == "diag":
                lg.append(G.LGate(G.DIAG, (b, b2), G.rzz_diag(0.3)))
            elif kind == "dense2":
                lg.append(G.LGate(G.DENSE, (b, b2), G.rxx_mat(0.4)))
        prog = compile_program(lg, n, TileConfig(m=m, L=L, threads=thr), itemsize=B)
        assert prog.n_passes == 1
        dp = P.DeviceProgram(prog, dev, tdt)
        dp.run(st); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dp.run(st)
        e1.record(); torch.cuda.synchronize()
        res.append((cnt, e0.elapsed_time(e1) / 3))
    slope = (res[-1][1] - res[0][1]) / (res[-1][0] - res[0][0])
    print(f"{kind:7s} n={n} {dt} m={m} L={L} thr={thr}: " + " ".join(f"G={c}:{ms:.2f}ms" for c, ms in res) + f"  -> {slope:.3f} ms per extra sweep", flush=True)


if __name__ == "__main__":
    main()
