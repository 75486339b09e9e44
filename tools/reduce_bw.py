"""Bandwidth of the read-only kernels (norm, <Z> on all bits, Z-mask sums, probabilities, CDF + sampling) on one GPU.
usage: reduce_bw.py [n] [c128|c64]   (run under gpurun; prints one line per kernel, GB/s = algorithmic bytes / time)"""
from __future__ import annotations

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.circuits import hea_ops  # noqa: E402
from tyxonq_b200.fuse import fuse  # noqa: E402
from tyxonq_b200.gates import lower_op  # noqa: E402
import numpy as np  # noqa: E402


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    dt = sys.argv[2] if len(sys.argv) > 2 else "c128"
    tdt = torch.complex128 if dt == "c128" else torch.complex64
    B = 16 if dt == "c128" else 8
    dev = torch.device("cuda", 0)
    _lib.ensure_device(0)
    st = P.new_state(n, dtype=tdt, device=dev)
    ops = hea_ops(n, 1, np.random.default_rng(1).uniform(-3, 3, 2 * n))
    P.apply_gates(st, fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None]))
    dim = 1 << n
    masks = [(1 << q) | (1 << ((q + 1) % n)) for q in range(n)]
    mt = torch.tensor(masks, dtype=torch.int64, device=dev)
    u = torch.from_numpy(np.random.default_rng(2).random(8192)).to(dev)
    rows = [
        ("norm2", lambda: P.norm2(st), dim * B),
        ("expect_z_bits (n values)", lambda: P.expect_z_bits(st), dim * B),
        (f"expect_zmasks ({len(masks)} masks)", lambda: P.expect_zmasks(st, mt), dim * B * ((len(masks) + 31) // 32)),
        ("cdf_chunks + sample (8192 shots)", lambda: P.sample(st, u), dim * B),
    ]
    if n <= 30:
        rows.append(("probabilities", lambda: P.probabilities(st), dim * (B + 8)))
    for name, fn, nbytes in rows:
        ms = timed(fn)
        print(f"n={n} {dt} {name:36s} {ms:8.3f} ms  {nbytes / ms / 1e6:8.0f} GB/s", flush=True)


if __name__ == "__main__":
    main()
