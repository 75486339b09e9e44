#!/usr/bin/env python
"""bench.py -- gate throughput of the statevector hot path on B200 (BASELINE.json metric:
gates/s and HBM GB/s vs peak at 30-36 qubits), one JSON line on rank 0.

Workload (config 3 of BASELINE.json): the reference's hardware-efficient ansatz
(libs/circuits_library/blocks.py:14-56: h on all, then per layer a cx chain + rz,rx on every
qubit), depth 100, angles default_rng(1234).uniform(-pi, pi), complex128 (the reference's only
dtype), n = 30 + log2(N) qubits on N GPUs (weak scaling: 2^30 amplitudes = 16 GiB per GPU).
A "step" = |0..0> -> all gates -> <Z_q> for every qubit, state resident in HBM.
`value` = gates/s normalised to 2^30-amplitude sweeps: gates * 2^(n-30) / t (at N = 1 exactly gates/s).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


# ------------------------------------------------------------------------------------------
def measured_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu_index = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.gpu_index), "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self) -> None:
        assert self.proc and self.proc.stdout
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload(n: int, layers: int):
    from tyxonq_b200.circuits import hea_ops
    params = np.random.default_rng(1234).uniform(-np.pi, np.pi, 2 * layers * n)
    ops = hea_ops(n, layers, params)
    return ops + [("measure_z", q) for q in range(n)], len(ops)


# ------------------------------------------------------------------------------------------
def cpu_baseline(n: int, layers: int, dtype_name: str, budget_s: float = 15.0) -> dict:
    """The oracle's C/OpenMP port of the reference kernels (kind "port") on a bounded sample: the
    first G gates of the same circuit on the same 2^n state, all host threads."""
    import psutil
    from oracle import c_oracle as CO
    npdt = np.complex128 if dtype_name == "complex128" else np.complex64
    need = (1 << n) * np.dtype(npdt).itemsize
    n_cpu = n
    while need * 1.5 > psutil.virtual_memory().available and n_cpu > 20:
        n_cpu -= 1
        need //= 2
    ops, _ = workload(n_cpu, layers)
    gate_ops = [o for o in ops if o[0] != "measure_z"]
    psi = CO.new_state(n_cpu, npdt)
    t0 = time.perf_counter()
    CO.apply_ops(psi, n_cpu, gate_ops[:4])
    per = (time.perf_counter() - t0) / 4
    G = int(max(8, min(len(gate_ops) - 4, budget_s / max(per, 1e-6))))
    t0 = time.perf_counter()
    done = CO.apply_ops(psi, n_cpu, gate_ops[4:4 + G])
    dt = time.perf_counter() - t0
    del psi
    scale = 2.0 ** (n_cpu - 30)
    return {"value": done * scale / dt, "unit": "gates/s", "cores": CO.max_threads(), "kind": "port",
            "sample": f"gates 5..{4 + done} of the same HEA circuit on a 2^{n_cpu} {dtype_name} state, in place, OpenMP "
                      f"({dt:.1f} s); numpy-einsum reference path is single-threaded and stops at n=22 (SURVEY fact 5)",
            "ms_per_gate": 1e3 * dt / done}


def reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n or (30 + int(math.log2(args.gpus)))
    vals = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(n, args.layers, args.dtype, budget_s=max(2.0, 60.0 / (args.warmup + args.steps)))
        if i >= args.warmup:
            vals.append(base)
    v = float(np.mean([b["value"] for b in vals]))
    ms = float(np.mean([b["ms_per_gate"] for b in vals]))
    base["value"] = v
    line = {"impl": "reference", "metric": "gates_per_s", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"hea{n}_depth{args.layers}_{args.dtype}", "n_qubits": n, "layers": args.layers,
                       "step": "bounded sample of the circuit's gates (see cpu_baseline.sample); ms_per_step = ms per gate"},
            "cpu_baseline": base,
            "e2e": {"value": v, "unit": "gates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------
def own_arm(args) -> None:
    import torch
    import torch.distributed as dist
    from tyxonq_b200 import StatevectorEngine, _lib
    from tyxonq_b200 import program as P
    from tyxonq_b200.circuits import Circuit
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import compile_program, default_tile

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _lib.ensure_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.n or (30 + int(math.log2(world)))
    tdt = torch.complex128 if args.dtype == "complex128" else torch.complex64
    B = 16 if args.dtype == "complex128" else 8
    peaks = measured_peaks()

    ops, n_gates = workload(n, args.layers)
    if world > 1:
        from tyxonq_b200.sharded import ShardedBench
        sb = ShardedBench(n, ops, tdt, dev)
        step, info = sb.step, sb.info
        n_local = sb.n_local
    else:
        t0 = time.perf_counter()
        lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
        prog = compile_program(lg, n, default_tile(n, B, 1), itemsize=B)
        plan_ms = 1e3 * (time.perf_counter() - t0)
        dp = P.DeviceProgram(prog, dev, tdt)
        state = torch.empty(1 << n, dtype=tdt, device=dev)
        pass_ev: list = []

        def step(timed: bool = False):
            ptr, n_, b_, dt_, stream = P._prep(state)
            _lib.check(_lib.load().tqb_init_basis(ptr, n_, 1, dt_, 0, 0, stream))
            if timed:
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record()
            dp.run(state)
            if timed:
                e1.record()
                pass_ev.append((e0, e1))
            return P.expect_z_bits(state)

        from tyxonq_b200.planner import fp_ops_per_amplitude
        info = {"passes": prog.n_passes, "gates_per_pass": n_gates / prog.n_passes, "fused_gate_sweeps": len(prog.gates), "plan_ms": plan_ms,
                "fp_ops_per_amplitude": fp_ops_per_amplitude(prog),
                "tile_m": prog.tile.m, "tile_L": prog.tile.L, "threads": prog.tile.threads, "swaps": 0}
        n_local = n

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        z = step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _lib.launch_count()
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        z = step(True)
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else {}
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    norm = 2.0 ** (n - 30)
    value = n_gates * norm / (ms_per_step * 1e-3)

    # roofline of the dominant kernel (tile_pass_lean_kernel: every pass of this workload is lean-eligible): algorithmic bytes 2 * 2^n_local * B per launch
    if world == 1:
        pass_ms = sum(a.elapsed_time(b) for a, b in pass_ev) / len(pass_ev)
        per_launch_ms = pass_ms / info["passes"]
    else:
        per_launch_ms = sb.pass_ms_per_launch()
        pass_ms = per_launch_ms * info["passes"]
    alg_bytes = 2.0 * (1 << n_local) * B
    achieved = alg_bytes / (per_launch_ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "kernel": "tile_pass_lean_kernel<double,2,128,GM=false,PAD=false|true>" if args.dtype == "complex128" else "tile_pass_lean_kernel<float,2,128,GM=false,PAD=false|true>", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
            "frac": achieved / peaks["hbm_gbs"], "peak_source": peaks["source"], "traffic": None,
            "alg_bytes_per_launch": alg_bytes, "launch_ms": per_launch_ms, "pass_share_of_step": pass_ms / ms_per_step}
    if "fp_ops_per_amplitude" in info:
        # the arithmetic side: the fused passes carry ~44 gates each, so the pass is bound by its in-tile compute phase.
        # Peak = measured DFMA / FFMA lane rate of this GPU (tools/micro/fp64_peak.cu: 32.6 / 69.3 TFLOP/s = 16.3 / 34.6 T instr/s).
        lane_ops = info["fp_ops_per_amplitude"] * (1 << n_local)
        peak_t = 16.3 if args.dtype == "complex128" else 34.6
        ach_t = lane_ops / (pass_ms * 1e-3) / 1e12
        roof["compute"] = {"bound": "fp64 pipe" if args.dtype == "complex128" else "fp32 pipe", "achieved": ach_t, "peak": peak_t,
                           "unit": "T lane-instr/s (mul / add / fma)", "frac": ach_t / peak_t,
                           "fp_instr_per_amplitude_per_step": info["fp_ops_per_amplitude"],
                           "peak_source": "tools/micro/fp64_peak.cu on this pool's B200 (register-resident FMA chains, 64 warps/SM)"}
    prof = ROOT / "profiles" / "r01_tile_pass_traffic.json"
    if prof.exists():
        try:
            roof["traffic"] = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            pass

    # end-to-end through the reference-facing API: host op list -> StatevectorEngine.run -> host dict
    e2e = None
    if world == 1:
        eng = StatevectorEngine("numpy", device=dev, dtype=tdt)
        circ = Circuit(n, ops)
        del state
        torch.cuda.empty_cache()
        eng.run(circ, shots=0)
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(1, min(2, args.steps))):
            t0 = time.perf_counter()
            res = eng.run(circ, shots=0)
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        assert len(res["expectations"]) == n
        e2e = {"value": n_gates * norm / float(np.mean(ts)), "unit": "gates/s", "h2d_bytes_per_step": int(eng.last_h2d_bytes),
               "d2h_bytes_per_step": int(eng.last_d2h_bytes), "ms_per_step": 1e3 * float(np.mean(ts)),
               "api": "StatevectorEngine.run(Circuit(n, ops), shots=0): plan + upload + passes + <Z_q> + D2H",
               "expz_checksum": float(sum(res["expectations"].values()))}
    else:
        e2e = sb.e2e(args)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    line = {"metric": "gates_per_s", "value": value, "unit": "gates/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"hea{n}_depth{args.layers}_{args.dtype}", "n_qubits": n, "layers": args.layers,
                       "gates": n_gates, "amplitudes_per_gpu": 1 << n_local, "state_gib_per_gpu": (1 << n_local) * B / 2 ** 30,
                       "l2": "state (>= 8 GiB) is far larger than the 126 MB L2, no flush needed",
                       "value_definition": "gates * 2^(n-30) / s", **info},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof,
            "hbm_gbps_per_gpu": achieved, "expz_checksum": float(z.sum().item())}
    if world > 1:
        bd = sb.breakdown
        # combined roofline: local passes at the measured HBM peak + exchanges at the measured NVLink peer peak, no overlap
        ideal_ms = info["passes"] * alg_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3 + \
            bd["exchange_bytes_per_gpu_each_way"] / (bd["nvlink_peak_gbps"] * 1e9) * 1e3
        bd["combined_roofline_ms"] = ideal_ms
        bd["combined_roofline_frac"] = ideal_ms / ms_per_step
        line["nvlink"] = bd
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(n, args.layers, args.dtype)
    if world == 1 and not args.no_extras:
        try:
            line["extras"] = extras(dev)
        except Exception as exc:  # extras never invalidate the headline
            line["extras"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def extras(dev) -> dict:
    """Secondary numbers of BASELINE.json's metric: complex64 gate throughput, VQE evals/s."""
    import torch
    from tyxonq_b200 import _lib, ucc
    from tyxonq_b200 import program as P
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import compile_program, default_tile
    from tyxonq_b200.vqe import TFIMVqe
    out = {}
    peaks = measured_peaks()
    # complex64 sweep at n = 30
    n, layers = 30, 20
    ops, n_gates = workload(n, layers)
    lg = fuse([g for g in (lower_op(o, n, mode="run") for o in ops) if g is not None])
    prog = compile_program(lg, n, default_tile(n, 8, 1), itemsize=8)
    dp = P.DeviceProgram(prog, dev, torch.complex64)
    st = P.new_state(n, dtype=torch.complex64, device=dev)
    dp.run(st); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); dp.run(st); dp.run(st); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 2
    gbps = prog.n_passes * 2.0 * (1 << n) * 8 / (ms * 1e-3) / 1e9
    out["complex64_hea30_depth20"] = {"gates_per_s": n_gates / (ms * 1e-3), "passes": prog.n_passes, "hbm_gbps": gbps,
                                      "frac_of_peak": gbps / peaks["hbm_gbs"]}
    del st, dp
    torch.cuda.empty_cache()
    # H2O-shaped UCCSD energy + adjoint gradient (config 2)
    i1, i2 = ucc.random_integral(7, 2077)
    ex_ops, pids = ucc.uccsd_ex_ops(5, 2)
    sv = ucc.UCCStatevector(14, (5, 5), ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=dev)
    np.random.seed(2077)
    p = np.random.rand(75) - 0.5
    sv.energy_and_grad(p)
    sv.energy_and_grad(p)
    t0 = time.perf_counter()
    reps = 50
    for _ in range(reps):
        e, g = sv.energy_and_grad(p)
    out["ucc_h2o_shape_energy_grad"] = {"evals_per_s": reps / (time.perf_counter() - t0), "energy": e, "n_params": 75,
                                        "excitations": 140, "pauli_terms": sv.ham.n_terms, "dtype": "complex128"}
    v = TFIMVqe(10, 1, device=dev)
    p = np.random.default_rng(0).normal(size=(2, 10))
    v.energy_and_grad(p)
    v.energy_and_grad(p)
    t0 = time.perf_counter()
    reps = 200
    for _ in range(reps):
        e, g = v.energy_and_grad(p)
    out["tfim10_energy_grad"] = {"evals_per_s": reps / (time.perf_counter() - t0), "energy": e,
                                 "note": "examples/vqetfim_benchmark.py ansatz + Hamiltonian, adjoint gradient, one CUDA graph per evaluation"}
    del sv, v
    torch.cuda.empty_cache()
    # config 5: 1024 parameter sets of the 20-qubit HWE-RY ansatz (L = 4), Heisenberg Pauli sum, 8192 shots per state
    from tyxonq_b200 import PauliSum
    from tyxonq_b200.batched import BatchedAnsatz
    nq, L, Bn, shots = 20, 4, 1024, 8192
    params = np.random.default_rng(7).random((Bn, (L + 1) * nq))
    terms = []
    for i in range(nq - 1):
        for c in ("Z", "X", "Y"):
            terms.append((1.0, [(c, i), (c, i + 1)]))
    ham = PauliSum.from_pauli_list(nq, terms)
    ba = BatchedAnsatz(nq, L, Bn, device=dev, dtype=torch.complex64)
    u_host = torch.from_numpy(np.random.default_rng(99).random((Bn, shots))).pin_memory()
    ba.run(params); torch.cuda.synchronize()
    t0 = time.perf_counter(); ba.run(params); torch.cuda.synchronize(); t_state = time.perf_counter() - t0
    ev = ba.expvals(ham); torch.cuda.synchronize()
    t0 = time.perf_counter(); ev = ba.expvals(ham); torch.cuda.synchronize(); t_ev = time.perf_counter() - t0
    idx = ba.sample(u_host.to(dev)); torch.cuda.synchronize()
    t0 = time.perf_counter(); idx = ba.sample(u_host.to(dev, non_blocking=True)); torch.cuda.synchronize(); t_s = time.perf_counter() - t0
    gates_per_state = nq * (L + 1) + (nq - 1) * L
    out["batched_hwe20_x1024_complex64"] = {
        "states_per_s": Bn / t_state, "gates_per_s_20q": Bn * gates_per_state / t_state, "passes": ba.passes,
        "state_hbm_gbps": ba.passes * 2.0 * Bn * (1 << nq) * 8 / t_state / 1e9,
        "expvals_per_s": Bn / t_ev, "pauli_terms": ham.n_terms, "xmask_groups": ham.n_groups,
        "shots_per_s": Bn * shots / t_s, "shots_per_state": shots,
        "energy_mean": float(ev.mean().item()), "idx_checksum": int(idx.sum().item())}
    return out


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n", type=int, default=0, help="qubits (default 30 + log2(gpus))")
    ap.add_argument("--layers", type=int, default=100)
    ap.add_argument("--dtype", default="complex128", choices=["complex128", "complex64"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
