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


def _program(name: str):
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import compile_program, default_tile
    bp = batched_program(name)
    if bp is not None:
        return bp
    n, ops, itemsize = shapes(name)
    lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
    return compile_program(lg, n, default_tile(n, itemsize), itemsize=itemsize), itemsize


def _compile_passes(name: str, indices) -> tuple:
    """Worker: NVRTC-compile the listed passes of a workload (a process of its own: the library is loaded per process)."""
    from tyxonq_b200 import _lib
    lib = _lib.load()
    prog, itemsize = _program(name)
    passes, gates = np.ascontiguousarray(prog.passes), np.ascontiguousarray(prog.gates)
    ok = 0
    for pi in indices:
        one = np.ascontiguousarray(passes[pi:pi + 1])
        ok += lib.tqb_spec_compile(one.ctypes.data, gates.ctypes.data, 1 if itemsize == 16 else 0) > 0
    return ok, _lib.jit_stats()["compiles"]


def prewarm(names, workers: int = 0) -> int:
    """Compile every pass shape of the named workloads into the disk cache.  ``workers`` > 1: the DISTINCT shapes (by
    generated header) are spread over that many processes (NVRTC takes 0.3-3 s per shape)."""
    import ctypes as C
    from tyxonq_b200 import _lib
    lib = _lib.load()
    done = 0
    jobs = []
    for name in names:
        prog, itemsize = _program(name)
        passes, gates = np.ascontiguousarray(prog.passes), np.ascontiguousarray(prog.gates)
        t0 = time.time()
        if workers > 1:   # one representative pass per distinct generated header
            buf = C.create_string_buffer(1 << 16)
            seen, reps, bad = set(), [], 0
            for pi in range(len(passes)):
                one = np.ascontiguousarray(passes[pi:pi + 1])
                r = lib.tqb_spec_source(one.ctypes.data, gates.ctypes.data, 1 if itemsize == 16 else 0, 0, buf, len(buf))
                if r < 0:
                    bad += 1
                elif buf.value not in seen:
                    seen.add(buf.value)
                    reps.append(pi)
            jobs.append((name, reps, len(passes), bad))
            continue
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
    if jobs:
        from concurrent.futures import ProcessPoolExecutor
        import multiprocessing as mp
        t0 = time.time()
        tasks = []
        for name, reps, _, _ in jobs:
            k = max(1, min(workers, len(reps)))
            tasks += [(name, reps[i::k]) for i in range(k)]
        with ProcessPoolExecutor(max_workers=workers, mp_context=mp.get_context("spawn")) as ex:
            res = list(ex.map(_compile_passes, [t[0] for t in tasks], [t[1] for t in tasks]))
        for name, reps, npass, bad in jobs:
            ok = sum(r[0] for r, t in zip(res, tasks) if t[0] == name)
            print(f"{name}: {npass} passes, {len(reps)} shapes, {ok} compiled or cached ({bad} passes generic)", flush=True)
            done += ok
        print(f"prewarm: {sum(len(j[1]) for j in jobs)} shapes in {time.time() - t0:.1f} s on {workers} processes", flush=True)
    return done


if __name__ == "__main__":
    args = sys.argv[1:]
    nw = 0
    if args and args[0].startswith("-j"):
        nw = int(args[0][2:] or 8)
        args = args[1:]
    prewarm(args or ["smoke", "hea30"], workers=nw)
