"""Config 4 of BASELINE.json: TFIM Trotter evolution (J = h = 1, t = 1, `steps` Trotter steps, the op list of
libs/circuits_library/trotter_circuit.py) on a state sharded by global qubits over all ranks.
usage: torchrun --nproc-per-node N tools/trotter_sharded.py [n] [steps] [c64|c128] [reps]
Prints one JSON line on rank 0: gates/s, HBM GB/s/GPU, NVLink GB/s/GPU, exchanges, combined-roofline fraction,
and two size-independent checks (norm = 1, <Z_q> = <Z_{n-1-q}> by the chain's reflection symmetry)."""
from __future__ import annotations

import json
import math
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from tyxonq_b200 import _lib  # noqa: E402
from tyxonq_b200 import program as P  # noqa: E402
from tyxonq_b200.circuits import tfim_terms, trotter_ops  # noqa: E402
from tyxonq_b200.sharded import ShardedBench  # noqa: E402


def main():
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 33 + int(math.log2(world))
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    dt = sys.argv[3] if len(sys.argv) > 3 else "c64"
    reps = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    tdt = torch.complex64 if dt == "c64" else torch.complex128
    B = 8 if dt == "c64" else 16
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    _lib.ensure_device(lr)
    dist.init_process_group("nccl", device_id=dev)
    ops = trotter_ops(*tfim_terms(n, 1.0, 1.0), 1.0, steps)
    n_gates = len([o for o in ops if o[0] != "measure_z"])
    sb = ShardedBench(n, ops, tdt, dev)
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
    zq = z.cpu().numpy()[::-1]  # index bit n-1-q -> qubit q
    if rank == 0:
        bd = sb.breakdown
        alg = 2.0 * (1 << sb.n_local) * B
        hbm_peak = 6552.6
        p = Path(__file__).resolve().parents[1] / "MEASURED_PEAKS.json"
        if p.exists():
            hbm_peak = float(json.loads(p.read_text())["hbm_gbs"])
        ideal = sb.info["passes"] * alg / (hbm_peak * 1e9) * 1e3 + bd["exchange_bytes_per_gpu_each_way"] / (770e9) * 1e3
        print(json.dumps({"workload": f"tfim_trotter{n}_steps{steps}_{dt}", "n_gpus": world, "n_qubits": n, "n_local": sb.n_local,
                          "gates": n_gates, "ms_per_step": ms, "gates_per_s": n_gates / (ms * 1e-3),
                          "passes": sb.info["passes"], "gates_per_pass": n_gates / sb.info["passes"],
                          "hbm_gbps_per_gpu": alg / (per_launch * 1e-3) / 1e9, "hbm_frac": alg / (per_launch * 1e-3) / 1e9 / hbm_peak,
                          "exchanges": bd["exchanges"], "exchange_ms": bd["exchange_ms"], "pass_ms": bd["pass_ms"],
                          "nvlink_gbps_per_gpu_each_way": bd["nvlink_gbps_per_gpu_each_way"],
                          "nvlink_frac_of_770": (bd["nvlink_gbps_per_gpu_each_way"] or 0) / 770.0,
                          "combined_roofline_ms": ideal, "combined_roofline_frac": ideal / ms,
                          "norm_minus_1": float(nrm[0]) - 1.0, "z_reflection_err": float(np.abs(zq - zq[::-1]).max()),
                          "z_first4": zq[:4].tolist(), "state_gib_per_gpu": (1 << sb.n_local) * B / 2 ** 30}), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
