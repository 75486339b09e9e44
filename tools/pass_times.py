"""Per-pass device time of one workload on one GPU: full / gates skipped (staging only) / staging skipped (gates only),
next to what the pass contains.  usage: python tools/pass_times.py hea30|hea30c64|qaoa30|trotter [max_passes [first_pass [full_only]]]"""
from __future__ import annotations

import collections
import os
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
from prewarm_jit import shapes  # noqa: E402

from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.fuse import fuse  # noqa: E402
from tyxonq_b200.gates import lower_op  # noqa: E402
from tyxonq_b200.planner import compile_program, default_tile  # noqa: E402

KIND = {0: "dense", 1: "diag", 2: "pair", 3: "swap", 4: "mux", 5: "chain"}


def main() -> None:
    name = sys.argv[1]
    limit = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    first = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    full_only = len(sys.argv) > 4
    n, ops, itemsize = shapes(name)
    dt = torch.complex128 if itemsize == 16 else torch.complex64
    lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
    prog = compile_program(lg, n, default_tile(n, itemsize), itemsize=itemsize)
    lib = _lib.load()
    lib.tqb_set_jit(2)
    dev = torch.device("cuda:0")
    state = P.new_state(n, dtype=dt, device=dev)
    dp = P.DeviceProgram(prog, dev, dt)
    ptr, _, batch, dtc, stream = P._prep(state)
    t = prog.tile
    passes = np.ascontiguousarray(prog.passes)
    alg = 2.0 * (1 << n) * itemsize

    def run(pi: int) -> float:
        one = np.ascontiguousarray(passes[pi:pi + 1])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.tqb_run_passes2(ptr, n, batch, dtc, 0, one.ctypes.data, 1, dp.gates_dev.data_ptr(), dp._gates_host.ctypes.data,
                                       dp.mats_dev.data_ptr(), t.threads, t.ctas_per_sm, stream))
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    tot = collections.defaultdict(float)
    np_ = min(limit, len(passes) - first)
    print(f"{name}: n={n} itemsize={itemsize} passes={len(passes)} gates={prog.n_gates_in} ({prog.n_gates_in / len(passes):.1f}/pass), "
          f"floor {alg / 6557.8e9 * 1e3:.2f} ms/pass at the measured HBM peak")
    for pi in range(first, first + np_):
        ps = passes[pi]
        gs = prog.gates[int(ps["gate_begin"]):int(ps["gate_begin"]) + int(ps["n_gates"])]
        comp = collections.Counter(f"{KIND.get(int(g['kind']), int(g['kind']))}{int(g['k'])}" for g in gs)
        res = {}
        modes = [("full", 256), ("staging", 256 + 1), ("gates", 256 + 6)]
        if os.environ.get("TQB_PASS_TIMES_REP"):
            modes += [("rep", 256 + 8), ("rep_gates", 256 + 14)]
        for label, mode in modes:
            if full_only and label != "full":
                res[label] = 0.0
                continue
            lib.tqb_set_jit(mode)
            if not full_only:
                run(pi)
            res[label] = run(pi) if full_only else min(run(pi), run(pi))
            tot[label] += res[label]
        lib.tqb_set_jit(256)
        hb = [int(x) for x in ps["hb"][: int(ps["m"]) - int(ps["L"])]]
        extra = "".join(f" {k} {v:6.2f}" for k, v in res.items() if k.startswith("rep"))
        print(f"pass {pi:3d} full {res['full']:6.2f} staging {res['staging']:6.2f} gates {res['gates']:6.2f}{extra} ms  hb={hb} "
              f"{dict(comp)}", flush=True)
    print(f"sum over {np_} passes: full {tot['full']:.1f} staging {tot['staging']:.1f} gates {tot['gates']:.1f} ms; "
          f"HBM fraction {np_ * alg / (tot['full'] * 1e-3) / 6557.8e9:.3f}; jit {_lib.jit_stats()}")


if __name__ == "__main__":
    main()
