"""One gate per pass, repeated: the pure streaming cost of the tile pass (run under gpurun / ncu).
usage: base_pass.py n dtype(c128|c64) m L threads qubit reps [gate]"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.gates import lower_op  # noqa: E402
from tyxonq_b200.planner import TileConfig, compile_program  # noqa: E402


def main():
    n, dt, m, L, thr, q, reps = int(sys.argv[1]), sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5]), int(sys.argv[6]), int(sys.argv[7])
    gate = sys.argv[8] if len(sys.argv) > 8 else "h"
    tdt = torch.complex128 if dt == "c128" else torch.complex64
    B = 16 if dt == "c128" else 8
    dev = torch.device("cuda", 0)
    _lib.ensure_device(0)
    st = P.new_state(n, dtype=tdt, device=dev)
    op = (gate, q) if gate in ("h", "x", "s") else ((gate, q, 0.3) if gate in ("rx", "rz") else (gate, q, (q + 1) % n))
    prog = compile_program([lower_op(op, n, mode="run")], n, TileConfig(m=m, L=L, threads=thr), itemsize=B)
    dp = P.DeviceProgram(prog, dev, tdt)
    dp.run(st)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dp.run(st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"n={n} {dt} m={m} L={L} thr={thr} q={q} gate={gate}: {ms:.3f} ms/pass, {2.0 * (1 << n) * B / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
