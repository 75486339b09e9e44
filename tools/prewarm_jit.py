"""Compile the specialised pass kernels of the standard workloads into the on-disk cubin cache (tyxonq_b200/jit_cache/)
with NVRTC -- needs no GPU.  The cache is git-ignored but ships to the GPU box with the snapshot, so the first step of a
bench or smoke run finds its kernels ready instead of compiling them in the background.
usage: python tools/prewarm_jit.py [hea30] [hea30c64] [qaoa30] [smoke] [trotter] [hwe20b]"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def shapes(name: str):
    from tyxonq_b200.circuits import hea_ops, qaoa_ring_ops, tfim_terms, trotter_ops
    rng = np.random.default_rng(1234)
    if name == "hea30":
        return 30, hea_ops(30, 100, rng.uniform(-np.pi, np.pi, 6000)), 16
    if name == "hea30c64":
        return 30, hea_ops(30, 100, rng.uniform(-np.pi, np.pi, 6000)), 8
    if name == "qaoa30":
        return 30, qaoa_ring_ops(30, 100, rng.uniform(-np.pi, np.pi, 200)), 16
    if name == "smoke":
        return 20, hea_ops(20, 2, np.random.default_rng(3).uniform(-np.pi, np.pi, 80)), 16
    if name == "trotter":
        return 33, [o for o in trotter_ops(*tfim_terms(33, 1.0, 1.0), 1.0, 10) if o[0] != "measure_z"], 8
    raise SystemExit(f"unknown workload {name}")


def batched_program(name: str):
    """Programs with one matrix set per batch member (config 5: 1024 parameter sets of the 20-qubit HWE-RY ansatz)."""
    from tyxonq_b200.batched import hwe_ry_gates
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.planner import compile_program, default_tile
    if name == "hwe20b":
        nq, L, Bn, itemsize = 20, 4, 1024, 8
        params = np.random.default_rng(7).random((Bn, (L + 1) * nq))
        return compile_program(fuse(hwe_ry_gates(nq, L, params)), nq, default_tile(nq, itemsize, Bn), batch_mats=Bn, itemsize=itemsize), itemsize
    return None


def prewarm(names) -> int:
    from tyxonq_b200 import _lib
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import compile_program, default_tile
    lib = _lib.load()
    done = 0
    for name in names:
        bp = batched_program(name)
        if bp is not None:
            prog, itemsize = bp
        else:
            n, ops, itemsize = shapes(name)
            lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
            prog = compile_program(lg, n, default_tile(n, itemsize), itemsize=itemsize)
        passes, gates = np.ascontiguousarray(prog.passes), np.ascontiguousarray(prog.gates)
        t0 = time.time()
        ok = bad = 0
        for pi in range(len(passes)):
            one = np.ascontiguousarray(passes[pi:pi + 1])
            rc = lib.tqb_spec_compile(one.ctypes.data, gates.ctypes.data, 1 if itemsize == 16 else 0)
            ok += rc > 0
            bad += rc <= 0
        st = _lib.jit_stats()
        print(f"{name}: {len(passes)} passes, {ok} specialised ({bad} generic), NVRTC compilations so far {st['compiles']}, "
              f"disk hits {st['disk_hits']}, {time.time() - t0:.1f} s", flush=True)
        done += ok
    return done


if __name__ == "__main__":
    prewarm(sys.argv[1:] or ["smoke", "hea30"])
