"""GPU tuning sweep for the tile pass (run under gpurun): base pass cost (one gate per pass) and
HEA throughput for a grid of tile shapes.  Prints one line per configuration."""
from __future__ import annotations

import itertools
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


def time_prog(prog, state, reps=2):
    dp = P.DeviceProgram(prog, state.device, state.dtype)
    dp.run(state)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        dp.run(state)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    layers = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    which = sys.argv[3] if len(sys.argv) > 3 else "both"
    dev = torch.device("cuda", 0)
    _lib.ensure_device(0)
    for tdt, B in ((torch.complex128, 16), (torch.complex64, 8)):
        if which == "c128" and B == 8 or which == "c64" and B == 16:
            continue
        state = P.new_state(n, dtype=tdt, device=dev)
        bytes_pass = 2.0 * (1 << n) * B
        # base cost: a single gate per pass, target bit low / mid / high
        for m, L, thr, cps in [(11 if B == 16 else 12, 5 if B == 16 else 6, 256, 0), (11 if B == 16 else 12, 4 if B == 16 else 5, 256, 0),
                               (12 if B == 16 else 13, 5 if B == 16 else 6, 256, 0), (10 if B == 16 else 11, 5 if B == 16 else 6, 256, 0),
                               (11 if B == 16 else 12, 6 if B == 16 else 7, 256, 0), (11 if B == 16 else 12, 5 if B == 16 else 6, 128, 0),
                               (11 if B == 16 else 12, 5 if B == 16 else 6, 512, 0), (11 if B == 16 else 12, 5 if B == 16 else 6, 256, 2),
                               (11 if B == 16 else 12, 5 if B == 16 else 6, 256, 3)]:
            for q in (n - 1, n // 2, 0):
                g = lower_op(("h", q), n, mode="run")
                prog = compile_program([g], n, TileConfig(m=m, L=L, threads=thr, ctas_per_sm=cps), itemsize=B)
                ms = time_prog(prog, state, 3)
                print(f"BASE B={B} m={m} L={L} thr={thr} cps={cps} qubit={q} ms={ms:.3f} GBps={bytes_pass / ms / 1e6:.0f}", flush=True)
        params = np.random.default_rng(1234).uniform(-np.pi, np.pi, 2 * layers * n)
        ops = hea_ops(n, layers, params)
        lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
        for m, L, thr, cps in itertools.product((10, 11, 12, 13), (4, 5, 6), (128, 256, 512), (0,)):
            if B == 8:
                m += 1; L += 1
            if (B << m) > 128 * 1024:
                continue
            prog = compile_program(lg, n, TileConfig(m=m, L=L, threads=thr, ctas_per_sm=cps), itemsize=B)
            ms = time_prog(prog, state, 1)
            print(f"HEA B={B} m={m} L={L} thr={thr} cps={cps} passes={prog.n_passes} ms={ms:.1f} ms/pass={ms / prog.n_passes:.2f} "
                  f"gates/s={len(ops) / ms * 1e3:.0f} GBps={prog.n_passes * bytes_pass / ms / 1e6:.0f}", flush=True)
        del state
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
