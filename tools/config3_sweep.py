"""Config 3 of BASELINE.json on one GPU: HEA and QAOA (ring), depth 100, n swept, complex128 and complex64.
usage: config3_sweep.py [depth] [c128 sizes, comma separated] [c64 sizes]     (run under gpurun)
One JSON line per (circuit, dtype, n): gates/s, passes, ms per pass, GB/s of algorithmic traffic and its fraction of
the measured HBM peak (MEASURED_PEAKS.json, else 6552.6 GB/s).  Timed with CUDA events after one warm-up run."""
from __future__ import annotations

import json
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.circuits import hea_ops, qaoa_ring_ops  # noqa: E402
from tyxonq_b200.fuse import fuse  # noqa: E402
from tyxonq_b200.gates import lower_op  # noqa: E402
from tyxonq_b200.planner import compile_program, default_tile  # noqa: E402


def main():
    depth = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    s128 = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "26,28,30").split(",") if x]
    s64 = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "26,28,30,32,33").split(",") if x]
    dev = torch.device("cuda", 0)
    _lib.ensure_device(0)
    peak = 6552.6
    pk = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
    if pk.exists():
        peak = float(json.loads(pk.read_text()).get("hbm_gbs", peak))
    for dt, sizes in (("c128", s128), ("c64", s64)):
        tdt = torch.complex128 if dt == "c128" else torch.complex64
        B = 16 if dt == "c128" else 8
        for n in sizes:
            st = P.new_state(n, dtype=tdt, device=dev)
            rng = np.random.default_rng(1234)
            for name, ops in (("hea", hea_ops(n, depth, rng.uniform(-np.pi, np.pi, 2 * depth * n))),
                              ("qaoa_ring", qaoa_ring_ops(n, depth, rng.uniform(-np.pi, np.pi, 2 * depth)))):
                lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
                prog = compile_program(lg, n, default_tile(n, B, 1), itemsize=B)
                dp = P.DeviceProgram(prog, dev, tdt)
                dp.run(st); torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); dp.run(st); e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1)
                gbs = prog.n_passes * 2.0 * (1 << n) * B / ms / 1e6
                print(json.dumps({"circuit": name, "dtype": dt, "n": n, "depth": depth, "gates": len(ops), "fused_sweeps": len(lg),
                                  "passes": prog.n_passes, "ms": round(ms, 2), "ms_per_pass": round(ms / prog.n_passes, 3),
                                  "gates_per_s": round(len(ops) / ms * 1e3, 1), "gates_per_pass": round(len(ops) / prog.n_passes, 1),
                                  "hbm_gbs": round(gbs, 1), "hbm_frac": round(gbs / peak, 3), "norm2": float(P.norm2(st)[0])}), flush=True)
            del st
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
