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
def _host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:  # noqa: BLE001
        return os.cpu_count() or 1


class CpuPort:
    """The oracle's C/OpenMP port of the reference kernels on ONE resident 2^n state (allocated once), all host threads
    (torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly)."""

    def __init__(self, n: int, layers: int, dtype_name: str) -> None:
        import psutil
        from oracle import c_oracle as CO
        self.CO = CO
        self.threads = CO.set_threads(_host_threads())
        npdt = np.complex128 if dtype_name == "complex128" else np.complex64
        n_cpu = min(n, 30)
        while (1 << n_cpu) * np.dtype(npdt).itemsize * 1.3 > psutil.virtual_memory().available and n_cpu > 20:
            n_cpu -= 1
        self.n_cpu, self.dtype_name = n_cpu, dtype_name
        ops, _ = workload(n_cpu, layers)
        self.gates = [o for o in ops if o[0] != "measure_z"]
        self.psi = CO.new_state(n_cpu, npdt)
        self.pos = 0

    def run(self, g: int) -> float:
        """Apply the next g gates of the circuit (wrapping around); returns seconds."""
        sel = [self.gates[(self.pos + i) % len(self.gates)] for i in range(g)]
        self.pos = (self.pos + g) % len(self.gates)
        t0 = time.perf_counter()
        self.CO.apply_ops(self.psi, self.n_cpu, sel)
        return time.perf_counter() - t0


def cpu_baseline(n: int, layers: int, dtype_name: str, budget_s: float = 15.0) -> dict:
    """kind "port": the first gates of the same circuit on a 2^min(n,30) state with the oracle's C/OpenMP loops, for about
    ``budget_s`` seconds.  The reference's own numpy path (single-threaded einsum, stops at n = 22) is timed beside it at
    n = 20 when baseline/_ref is present (``reference_numpy``)."""
    port = CpuPort(n, layers, dtype_name)
    per = port.run(4) / 4
    G = int(max(8, min(len(port.gates), budget_s / max(per, 1e-6))))
    dt = port.run(G)
    scale = 2.0 ** (port.n_cpu - 30)
    out = {"value": G * scale / dt, "unit": "gates/s", "cores": port.threads, "kind": "port",
           "sample": f"{G} consecutive gates of the same HEA circuit on a 2^{port.n_cpu} {dtype_name} state, in place, "
                     f"C/OpenMP port of the reference kernels ({dt:.1f} s)",
           "ms_per_gate": 1e3 * dt / G}
    del port
    try:
        from tools import ref_baselines as RB
        if RB.reference_available():
            out["reference_numpy"] = RB.hea_gates_per_s_reference(20, 1, 2)
    except Exception as exc:  # noqa: BLE001 -- the extra line never invalidates the baseline
        out["reference_numpy"] = {"error": repr(exc)}
    return out


def reference_arm(args) -> None:
    """`--impl reference`: the CPU implementation of the path on this box's host cores, same metric / config / unit.
    Under torchrun only rank 0 works; the state is allocated once; every step is a bounded sample of G gates so that
    W + K steps end within about a minute."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = args.n or (30 + int(math.log2(args.gpus)))
    port = CpuPort(n, args.layers, args.dtype)
    per = port.run(4) / 4
    total_steps = args.warmup + args.steps
    G = int(max(2, min(len(port.gates), 50.0 / total_steps / max(per, 1e-6))))
    ts = []
    for i in range(total_steps):
        dt = port.run(G)
        if i >= args.warmup:
            ts.append(dt)
    dt = float(np.mean(ts))
    scale = 2.0 ** (port.n_cpu - 30)
    v = G * scale / dt
    base = {"value": v, "unit": "gates/s", "cores": port.threads, "kind": "port",
            "sample": f"each step = {G} consecutive gates of the same HEA circuit on a 2^{port.n_cpu} {args.dtype} state "
                      f"(resident, in place), C/OpenMP port of the reference kernels, {port.threads} threads; value normalised "
                      f"by 2^({port.n_cpu}-30)", "ms_per_gate": 1e3 * dt / G}
    del port
    try:
        from tools import ref_baselines as RB
        if RB.reference_available():
            base["reference_numpy"] = RB.hea_gates_per_s_reference(20, 1, 2)
    except Exception as exc:  # noqa: BLE001
        base["reference_numpy"] = {"error": repr(exc)}
    line = {"impl": "reference", "metric": "gates_per_s", "value": v, "unit": "gates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": f"hea{n}_depth{args.layers}_{args.dtype}", "n_qubits": n, "layers": args.layers,
                       "value_definition": "gates * 2^(n-30) / s",
                       "step": f"bounded sample: {G} gates per step (see cpu_baseline.sample)"},
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
        # the global<->local exchange is one big send/recv per peer: give NCCL's point-to-point path enough channels to
        # fill NVLink (with 2 ranks the default left the link at 0.56 of the peer-copy peak, SCALE_r01)
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "32")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "32")
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

    for w in range(args.warmup):
        z = step()
        if w == 0:
            # the specialised pass kernels of this circuit compile on background threads while the generic kernel runs
            # the first step: wait for them once, so that the remaining warm-up and the timed steps run the steady state
            torch.cuda.synchronize()
            _lib.load().tqb_jit_wait()
    barrier()
    jit0 = _lib.jit_stats()
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
    jit1 = _lib.jit_stats()
    spec_launches = jit1["spec_launches"] - jit0["spec_launches"]
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
    kname = ("tqb_spec_pass (csrc/tqb_spec.cuh, NVRTC-specialised per pass shape; "
             f"{spec_launches} of the timed launches, the rest tile_pass_lean_kernel)")
    roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s",
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
    prof = ROOT / "profiles" / ("r02_spec_pass_traffic.json" if args.dtype == "complex128" else "r01_tile_pass_traffic.json")
    if prof.exists():
        try:
            roof["traffic"] = json.loads(prof.read_text()).get("dram_bytes_per_launch")
        except Exception:
            pass

    # end-to-end through the reference-facing API: host op list -> StatevectorEngine.run -> host dict
    e2e = None
    if world == 1:
        del state
        torch.cuda.empty_cache()
        e2e = e2e_single(n, ops, n_gates, norm, dev, tdt, args)
    else:
        e2e = sb.e2e(args)

    trotter = None
    if world > 1 and not args.no_extras:
        # config 4 of BASELINE.json rides in the same line: TFIM Trotter evolution, complex64, 33 + log2(N) qubits
        # (36 qubits = 64 GiB per GPU on 8 GPUs), sharded by global qubits -- every rank takes part
        try:
            sb_breakdown = sb.breakdown
            del sb
            torch.cuda.empty_cache()
            trotter = trotter_config4(world, rank, dev, peaks)
        except Exception as exc:  # noqa: BLE001 -- never invalidates the headline
            trotter = {"error": repr(exc)}
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
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "specialised_launches": int(spec_launches), "roofline": roof,
            "hbm_gbps_per_gpu": achieved, "expz_checksum": float(z.sum().item())}
    if world > 1:
        bd = sb_breakdown if trotter is not None else sb.breakdown
        # combined roofline: local passes at the measured HBM peak + exchanges at the measured NVLink peer peak, no overlap
        ideal_ms = info["passes"] * alg_bytes / (peaks["hbm_gbs"] * 1e9) * 1e3 + \
            bd["exchange_bytes_per_gpu_each_way"] / (bd["nvlink_peak_gbps"] * 1e9) * 1e3
        bd["combined_roofline_ms"] = ideal_ms
        bd["combined_roofline_frac"] = ideal_ms / ms_per_step
        line["nvlink"] = bd
        if trotter is not None:
            line["extras"] = {"tfim_trotter_complex64_sharded": trotter}
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(n, args.layers, args.dtype)
    if world == 1 and not args.no_extras:
        try:
            line["extras"] = extras(dev, quick=args.quick_extras)
        except Exception as exc:  # extras never invalidate the headline
            line["extras"] = {"error": repr(exc)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def e2e_single(n, ops, n_gates, norm, dev, tdt, args) -> dict:
    """The same step end to end from HOST data: op list -> plan -> pinned H2D of descriptors / matrices -> passes ->
    <Z_q> -> D2H.  With the reference installed in baseline/_ref the call is the reference's own public chain on a real
    tyxonq.Circuit -- Circuit.device(provider="simulator", device="statevector").run(shots=0) (core/ir/circuit.py:868-981
    -> devices/base.py:252 -> simulators/driver.py:86, compile stage included) -- with tyxonq_b200.install() routing it
    to the B200 engine; otherwise the engine's own run() on the stand-in Circuit."""
    import torch
    from tyxonq_b200 import StatevectorEngine
    from tyxonq_b200.circuits import Circuit
    api = None
    run = None
    if tdt == torch.complex128:
        try:
            from tools import ref_baselines as RB
            if RB.reference_available():
                tq = RB.import_reference()
                tq.set_backend("numpy")
                import tyxonq_b200
                tyxonq_b200.install()
                circ = tq.Circuit(n, ops=list(ops))

                def run():
                    r = circ.device(provider="simulator", device="statevector").run(shots=0)
                    r = r[0] if isinstance(r, list) else r
                    inner = r.get("result_meta")   # the facade nests the driver's result dict under result_meta
                    r = inner if isinstance(inner, dict) and "expectations" in inner else r
                    if r.get("error"):
                        raise RuntimeError(r["error"])
                    return r
                api = ("tyxonq.Circuit(n, ops).device(provider='simulator', device='statevector').run(shots=0) of the unmodified "
                       "reference in baseline/_ref with tyxonq_b200.install(): compile stage + driver + plan + H2D + passes + <Z_q> + D2H "
                       "(statevector / probabilities of the result stay on the device as LazyHostArray)")
        except Exception as exc:  # noqa: BLE001 -- fall back to the engine's own API, say why
            api = None
            run = None
            sys.stderr.write(f"[bench] reference chain unavailable for e2e: {exc!r}\n")
    if run is None:
        eng = StatevectorEngine("numpy", device=dev, dtype=tdt)
        circ = Circuit(n, ops)

        def run():
            return eng.run(circ, shots=0)
        api = "StatevectorEngine.run(Circuit(n, ops), shots=0): plan + upload + passes + <Z_q> + D2H"
    try:
        run()
        torch.cuda.synchronize()
        ts = []
        for _ in range(max(1, min(3, args.steps))):
            t0 = time.perf_counter()
            res = run()
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
    finally:
        # the CPU baselines that follow time the REFERENCE's own engine and kernels: undo the routing
        import tyxonq_b200
        if "tyxonq" in sys.modules:
            tyxonq_b200.uninstall()
    assert len(res["expectations"]) == n
    stats = StatevectorEngine.LAST
    return {"value": n_gates * norm / float(np.mean(ts)), "unit": "gates/s", "h2d_bytes_per_step": int(stats.get("h2d_bytes", 0)),
            "d2h_bytes_per_step": int(stats.get("d2h_bytes", 0)), "ms_per_step": 1e3 * float(np.mean(ts)), "api": api,
            "expz_checksum": float(sum(res["expectations"].values()))}


def trotter_config4(world: int, rank: int, dev, peaks: dict, steps: int = 10, reps: int = 3) -> dict:
    """Config 4: TFIM Trotter evolution (libs/circuits_library/trotter_circuit.py:8-122, J = h = 1, t = 1, 10 steps) on a
    complex64 state of 33 + log2(N) qubits sharded by global qubits.  Combined roofline = local passes at the measured HBM
    peak + exchanges at the measured NVLink peer peak (no overlap assumed)."""
    import torch
    import torch.distributed as dist
    from tyxonq_b200 import program as P
    from tyxonq_b200.circuits import tfim_terms, trotter_ops
    from tyxonq_b200.sharded import ShardedBench
    n = 33 + int(math.log2(world))
    ops = trotter_ops(*tfim_terms(n, 1.0, 1.0), 1.0, steps)
    n_gates = len([o for o in ops if o[0] != "measure_z"])
    sb = ShardedBench(n, ops, torch.complex64, dev)
    z = sb.step()
    torch.cuda.synchronize()
    from tyxonq_b200 import _lib
    _lib.load().tqb_jit_wait()
    z = sb.step()
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        z = sb.step()
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t[0])
    per_launch = sb.pass_ms_per_launch()
    nrm = P.norm2(sb.st.state)
    dist.all_reduce(nrm)
    zq = z.cpu().numpy()[::-1]
    bd = sb.breakdown
    alg = 2.0 * (1 << sb.n_local) * 8
    ideal = sb.info["passes"] * alg / (peaks["hbm_gbs"] * 1e9) * 1e3 + bd["exchange_bytes_per_gpu_each_way"] / (bd["nvlink_peak_gbps"] * 1e9) * 1e3
    out = {"workload": f"tfim_trotter{n}_steps{steps}_complex64", "n_qubits": n, "n_local": sb.n_local, "gates": n_gates,
           "ms_per_step": ms, "gates_per_s": n_gates / (ms * 1e-3), "value_2pow30": n_gates * 2.0 ** (n - 30) / (ms * 1e-3),
           "passes": sb.info["passes"], "gates_per_pass": n_gates / max(sb.info["passes"], 1), "repetitions": reps,
           "hbm_gbps_per_gpu": alg / (per_launch * 1e-3) / 1e9, "hbm_frac": alg / (per_launch * 1e-3) / 1e9 / peaks["hbm_gbs"],
           "exchanges": bd["exchanges"], "exchange_ms": bd["exchange_ms"], "pass_ms": bd["pass_ms"],
           "nvlink_gbps_per_gpu_each_way": bd["nvlink_gbps_per_gpu_each_way"],
           "combined_roofline_ms": ideal, "combined_roofline_frac": ideal / ms,
           "norm_minus_1": float(nrm[0]) - 1.0, "z_reflection_err": float(np.abs(zq - zq[::-1]).max()),
           "state_gib_per_gpu": (1 << sb.n_local) * 8 / 2 ** 30}
    del sb
    torch.cuda.empty_cache()
    return out


def _median_ms(fn, reps: int) -> float:
    """Median device time of fn() over reps repetitions (CUDA events on the current stream)."""
    import torch
    ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def _gate_sweep(name: str, ops, n: int, tdt, dev, reps: int = 5) -> dict:
    """One more circuit of config 3 through the same pass engine: gates/s, HBM GB/s of the passes and their fraction
    of the measured peak, with the CPU port beside it."""
    import torch
    from tyxonq_b200 import _lib
    from tyxonq_b200 import program as P
    from tyxonq_b200.fuse import fuse
    from tyxonq_b200.gates import lower_op
    from tyxonq_b200.planner import compile_program, default_tile
    peaks = measured_peaks()
    B = 16 if tdt == torch.complex128 else 8
    gate_ops = [o for o in ops if o[0] != "measure_z"]
    lg = fuse([g for g in (lower_op(o, n, mode="run") for o in gate_ops) if g is not None])
    prog = compile_program(lg, n, default_tile(n, B, 1), itemsize=B)
    dp = P.DeviceProgram(prog, dev, tdt)
    st = P.new_state(n, dtype=tdt, device=dev)
    dp.run(st); torch.cuda.synchronize()
    _lib.load().tqb_jit_wait()
    dp.run(st); torch.cuda.synchronize()
    j0 = _lib.jit_stats()["spec_launches"]
    ms = _median_ms(lambda: dp.run(st), reps)
    spec = (_lib.jit_stats()["spec_launches"] - j0) // reps
    gbps = prog.n_passes * 2.0 * (1 << n) * B / (ms * 1e-3) / 1e9
    del st, dp
    torch.cuda.empty_cache()
    out = {"gates_per_s": len(gate_ops) / (ms * 1e-3), "gates": len(gate_ops), "passes": prog.n_passes,
           "gates_per_pass": len(gate_ops) / prog.n_passes, "specialised_passes": int(spec), "ms": ms, "repetitions": reps,
           "roofline": {"bound": "hbm", "achieved": gbps, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbps / peaks["hbm_gbs"],
                        "alg_bytes_per_launch": 2.0 * (1 << n) * B}}
    try:
        from oracle import c_oracle as CO
        threads = CO.set_threads(_host_threads())
        npdt = np.complex128 if B == 16 else np.complex64
        psi = CO.new_state(n, npdt)
        CO.apply_ops(psi, n, gate_ops[:2])
        t0 = time.perf_counter()
        G = 0
        while time.perf_counter() - t0 < 4.0 and 2 + G < len(gate_ops):
            CO.apply_ops(psi, n, gate_ops[2 + G:2 + G + 4])
            G += 4
        dt = time.perf_counter() - t0
        del psi
        out["cpu_baseline"] = {"value": G / dt, "unit": "gates/s", "cores": threads, "kind": "port",
                               "sample": f"{G} consecutive gates of the same circuit on a 2^{n} state, C/OpenMP port ({dt:.1f} s)"}
    except Exception as exc:  # noqa: BLE001
        out["cpu_baseline"] = {"error": repr(exc)}
    return out


def extras(dev, quick: bool = False) -> dict:
    """The other configurations of BASELINE.json, each with its CPU baseline and a roofline fraction beside it."""
    import torch
    from tools import ref_baselines as RB
    from tyxonq_b200 import _lib, ucc
    from tyxonq_b200.circuits import qaoa_ring_ops
    from tyxonq_b200.vqe import TFIMVqe
    out = {}
    peaks = measured_peaks()
    have_ref = RB.reference_available()
    # ---- config 3, the other half: complex64 at full depth, QAOA ring at full depth (complex128)
    n, layers = 30, (20 if quick else 100)
    ops, _ = workload(n, layers)
    out[f"hea30_depth{layers}_complex64"] = _gate_sweep("hea", ops, n, torch.complex64, dev)
    qops = qaoa_ring_ops(n, layers, np.random.default_rng(1234).uniform(-np.pi, np.pi, 2 * layers))
    out[f"qaoa30_depth{layers}_complex128"] = _gate_sweep("qaoa", qops, n, torch.complex128, dev)
    # ---- config 2: H2O-shaped UCCSD energy + adjoint gradient
    i1, i2 = ucc.random_integral(7, 2077)
    ex_ops, pids = ucc.uccsd_ex_ops(5, 2)
    sv = ucc.UCCStatevector(14, (5, 5), ex_ops, pids, ucc.hamiltonian_from_integral(i1, i2), device=dev)
    np.random.seed(2077)
    p = np.random.rand(75) - 0.5
    sv.energy_and_grad(p)
    sv.energy_and_grad(p)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(20):
            e, g = sv.energy_and_grad(p)
        ts.append((time.perf_counter() - t0) / 20)
    dt = float(np.median(ts))
    # bytes an evaluation must move if the 256 KiB ket / bra stayed on chip: none -- the bound is the latency of 2 x 140
    # dependent pair rotations + H|psi>; reported as time per dependent step
    pbatch = np.random.default_rng(5).uniform(-0.5, 0.5, (128, 75))
    sv.energy_and_grad_batch(pbatch[:32], replicas=32)
    t0 = time.perf_counter()
    sv.energy_and_grad_batch(pbatch, replicas=32)
    t_batch = time.perf_counter() - t0
    out["ucc_h2o_shape_energy_grad"] = {"evals_per_s": 1.0 / dt, "batched_evals_per_s": 128 / t_batch,
                                        "batch": "128 parameter vectors on 32 concurrent replicas (own buffers, stream, CUDA graph and reduction-workspace slice each; the sweeps of a replica run in one CTA)",
                                        "energy": e, "n_params": 75, "excitations": 140,
                                        "pauli_terms": sv.ham.n_terms, "dtype": "complex128", "repetitions": "5 x 20",
                                        "us_per_dependent_step": 1e6 * dt / (2 * 140 + 2),
                                        "cpu_baseline": RB.ucc_h2o_energy_grad_port(2)}
    # ---- config 1: TFIM-10 VQE energy + gradient
    v = TFIMVqe(10, 1, device=dev)
    p = np.random.default_rng(0).normal(size=(2, 10))
    v.energy_and_grad(p)
    v.energy_and_grad(p)
    ts = []
    for _ in range(5):
        t0 = time.perf_counter()
        for _ in range(100):
            e, g = v.energy_and_grad(p)
        ts.append((time.perf_counter() - t0) / 100)
    dt = float(np.median(ts))
    pb = np.random.default_rng(1).normal(size=(1024, 20))
    v.energy_and_grad_batch(pb)
    tb = []
    for _ in range(5):
        t0 = time.perf_counter(); v.energy_and_grad_batch(pb); tb.append(time.perf_counter() - t0)
    tg = []
    for _ in range(3):
        t0 = time.perf_counter()
        for _ in range(50):
            v.energy_and_grad(p, resident=False)
        tg.append((time.perf_counter() - t0) / 50)
    out["tfim10_energy_grad"] = {"evals_per_s": 1.0 / dt, "energy": e, "repetitions": "5 x 100",
                                 "batched_evals_per_s": 1024 / float(np.median(tb)), "batch": 1024,
                                 "fused_pass_graph_path_evals_per_s": 1.0 / float(np.median(tg)),
                                 "note": "examples/vqetfim_benchmark.py ansatz + Hamiltonian, adjoint gradient: CTA-resident kernel (csrc/tqb_vqe.cu), one "
                                         "evaluation per call (host round trip included) and 1024 parameter vectors per launch; the round-1 path "
                                         "(fused passes, one CUDA graph per evaluation) beside it",
                                 "cpu_baseline": RB.tfim10_energy_grad_reference(5) if have_ref else {"error": "baseline/_ref missing"}}
    del sv, v
    torch.cuda.empty_cache()
    # ---- config 5: 1024 parameter sets of the 20-qubit HWE-RY ansatz (L = 4), Heisenberg Pauli sum, 8192 shots per state
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
    u_dev = torch.from_numpy(np.random.default_rng(99).random((Bn, shots))).pin_memory().to(dev)
    ba.run(params); torch.cuda.synchronize()
    _lib.load().tqb_jit_wait()     # the passes' specialised kernels (per-member matrices) are compiled / loaded from the cache
    ba.run(params); torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        t0 = time.perf_counter(); ba.run(params); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    t_state = float(np.median(ts))
    ev = ba.expvals(ham); torch.cuda.synchronize()
    t_ev = _median_ms(lambda: ba.expvals(ham), 5) * 1e-3
    ev = ba.expvals(ham)
    idx = ba.sample(u_dev); torch.cuda.synchronize()
    t_s = _median_ms(lambda: ba.sample(u_dev), 5) * 1e-3
    idx = ba.sample(u_dev)
    state_bytes = float(Bn) * (1 << nq) * 8
    gates_per_state = nq * (L + 1) + (nq - 1) * L
    ref5 = RB.batched_hwe20_reference(nq, L, shots, 1) if have_ref else {"error": "baseline/_ref missing"}
    out["batched_hwe20_x1024_complex64"] = {
        "states_per_s": Bn / t_state, "gates_per_s_20q": Bn * gates_per_state / t_state, "passes": ba.passes,
        "state_roofline": {"bound": "hbm", "achieved": ba.passes * 2.0 * state_bytes / t_state / 1e9, "peak": peaks["hbm_gbs"],
                           "unit": "GB/s", "frac": ba.passes * 2.0 * state_bytes / t_state / 1e9 / peaks["hbm_gbs"],
                           "note": "wall time of BatchedAnsatz.run: planning + upload of 1024 x 100 matrices + passes"},
        "expvals_per_s": Bn / t_ev, "pauli_terms": ham.n_terms, "xmask_groups": ham.n_groups,
        "expval_roofline": {"bound": "hbm", "achieved": state_bytes / t_ev / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": state_bytes / t_ev / 1e9 / peaks["hbm_gbs"], "note": "algorithmic bytes = ONE read of the 1024 states"},
        "shots_per_s": Bn * shots / t_s, "shots_per_state": shots,
        "sample_roofline": {"bound": "hbm", "achieved": state_bytes / t_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                            "frac": state_bytes / t_s / 1e9 / peaks["hbm_gbs"], "note": "algorithmic bytes = ONE read of the 1024 states"},
        "repetitions": 5, "energy_mean": float(ev.mean().item()), "idx_checksum": int(idx.sum().item()),
        "cpu_baseline": ref5}
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
    ap.add_argument("--quick-extras", action="store_true", help="config-3 extras at depth 20 instead of 100")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        own_arm(args)


if __name__ == "__main__":
    main()
