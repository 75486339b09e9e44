"""HEA throughput for explicit tile configs (run under gpurun).
usage: hea_cfg.py n layers dtype m:L:thr[:cps] [m:L:thr ...]"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.circuits import hea_ops  # noqa: E402
from tyxonq_b200.fuse import fuse  # noqa: E402
from tyxonq_b200.gates import lower_op  # noqa: E402
from tyxonq_b200.planner import TileConfig, compile_program  # noqa: E402


def main():
    n, layers, dt = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    import os
    MICRO = os.environ.get('TQB_MICRO', '1') == '1'
    tdt = torch.complex128 if dt == "c128" else torch.complex64
    B = 16 if dt == "c128" else 8
    dev = torch.device("cuda", 0)
    _lib.ensure_device(0)
    st = P.new_state(n, dtype=tdt, device=dev)
    params = np.random.default_rng(1234).uniform(-np.pi, np.pi, 2 * layers * n)
    ops = hea_ops(n, layers, params)
    lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
    for cfg in sys.argv[4:]:
        f = [int(x) for x in cfg.split(":")]
        m, L, thr = f[:3]
        cps = f[3] if len(f) > 3 else 0
        rl = int(os.environ.get("TQB_RROT", "4"))
        prog = compile_program(lg, n, TileConfig(m=m, L=L, threads=thr, ctas_per_sm=cps, rot_layers=rl), itemsize=B)
        dp = P.DeviceProgram(prog, dev, tdt)
        dp.run(st); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); dp.run(st); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        print(f"HEA n={n} {dt} m={m} L={L} thr={thr} cps={cps} passes={prog.n_passes} sweeps/pass={len(prog.gates) / prog.n_passes:.1f} ms={ms:.1f} "
              f"ms/pass={ms / prog.n_passes:.2f} gates/s={len(ops) / ms * 1e3:.0f} GBps={prog.n_passes * 2.0 * (1 << n) * B / ms / 1e6:.0f}", flush=True)


if __name__ == "__main__":
    main()
