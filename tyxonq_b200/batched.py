"""Batched parameter sets: B states of n qubits as one [B, 2^n] device array.

Workload of examples/vqe_batched_architecture_search.py (config 5 of BASELINE.json): the hardware-
efficient RY ansatz of libs/circuits_library/blocks.py:60-85 for B parameter sets at once, then a
matrix-free Pauli-sum expectation per state (instead of the dense 2^n x 2^n Hamiltonian of
kernels/pauli.py:74-87) and shots per state from host-supplied uniforms.  The batch index is just more
tile-index bits for the pass kernel; gates carry one matrix per batch member (tqb_gate.mat_bstride).
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _lib
from . import program as P
from .fuse import fuse
from .gates import LGate, dense_gate, swap_gate
from .pauli import PauliSum
from .planner import TileConfig, compile_program, default_tile


def hwe_ry_gates(n: int, layers: int, params: np.ndarray) -> List[LGate]:
    """build_hwe_ry_ops (blocks.py:60-85) for params [B, (layers+1)*n]: ry layers carry B matrices each."""
    p = np.asarray(params, dtype=np.float64).reshape(params.shape[0], layers + 1, n)
    c, s = np.cos(0.5 * p), np.sin(0.5 * p)
    ry = np.stack([c, -s, s, c], axis=-1).astype(np.complex128)  # [B, layer, qubit, 4]
    gates: List[LGate] = [dense_gate(ry[:, 0, q], [q], n, name="ry") for q in range(n)]
    for l in range(layers):
        gates += [swap_gate([q, q + 1], n, 0b10, 0b11, name="cx") for q in range(n - 1)]
        gates += [dense_gate(ry[:, l + 1, q], [q], n, name="ry") for q in range(n)]
    return gates


class HweRyRefill:
    """The matrix buffer of the batched HWE-RY program as a direct function of the parameters.

    The plan of ``compile_program(fuse(hwe_ry_gates(..)))`` depends only on (n, layers, tile); lowering B x 100 matrices through
    the generic pipeline costs more host time than the passes take on the device.  This class runs the pipeline ONCE on a
    two-member probe with distinct angles, reads off for every packed gate which parameter feeds which entry -- entries of
    DENSE / MUX / general CHAIN gates are +-cos(theta_k/2), +-sin(theta_k/2), 0 or 1; rotation-form chains (gates.chain_gate)
    carry a unit table and per layer (a, r) = (cos, +-sin) of one parameter -- checks the recipe on the probe's second
    member to 1e-13, and then refills the buffer of a B-member program with a handful of vectorised numpy operations.
    ``mats(params)`` returns None when a parameter set leaves the recipe's domain (a scaled layer that needs the c form,
    gates.ROT_T_MAX): the caller then lowers that call generically."""

    def __init__(self, n: int, layers: int, prog_b, batch: int, tile: TileConfig, itemsize: int) -> None:
        from .gates import CHAIN, ROT_T_MAX
        self.ok = False
        self.t_max = ROT_T_MAX
        self.batch = int(batch)
        npar = (layers + 1) * n
        rng = np.random.default_rng(20240229)
        probe = np.sort(rng.uniform(0.35, 1.25, 2 * npar)).reshape(npar, 2).T.copy()   # distinct angles, t form everywhere
        rng.shuffle(probe[0])
        rng.shuffle(probe[1])
        prog = compile_program(fuse(hwe_ry_gates(n, layers, probe)), n, tile, batch_mats=2, itemsize=itemsize)
        g2, gb = prog.gates, prog_b.gates
        if len(g2) != len(gb) or any(not np.array_equal(g2[f], gb[f]) for f in ("kind", "k", "bits", "off_a", "off_b")):
            return
        if not np.array_equal(g2["mat_bstride"], gb["mat_bstride"]):
            return
        c, s = np.cos(0.5 * probe), np.sin(0.5 * probe)
        self.total = int(prog_b.mats.size)
        self.plain: List[tuple] = []      # (offset, len, table index [len], sign [len])
        self.rot: dict = {}               # (R, tab) -> {"off": [...], "K": [[...]], "SG": [[...]]}
        self.const = np.zeros(self.total, dtype=np.complex128)   # entries that do not depend on the parameters
        tab0 = np.concatenate([[0.0, 1.0], c[0], s[0]])
        for gi in range(len(g2)):
            ln = int(g2[gi]["mat_bstride"])
            off2, offb = int(g2[gi]["mat_off"]), int(gb[gi]["mat_off"])
            if ln == 0:    # shared data (no batched gate upstream): copy it over as it is
                nxt = int(g2[gi + 1]["mat_off"]) if gi + 1 < len(g2) else int(prog.mats.size)
                self.const[offb:offb + (nxt - off2)] = prog.mats[off2:off2 + (nxt - off2)]
                continue
            d0 = prog.mats[off2:off2 + ln]
            if int(g2[gi]["kind"]) == CHAIN and int(g2[gi]["off_a"]) >= 4:
                ob = int(g2[gi]["off_b"])
                R = int(g2[gi]["k"])
                has_c = int(np.uint8(g2[gi]["bits"][R])) != 127
                tab = (1 << R) * (2 if has_c else 1)
                if not (ob & 128) or (ob & 3) or ((ob >> 8) & 0x1F) != 8 or ln != tab + R or ((int(g2[gi]["off_a"]) >> 1) & 1) != 1:
                    return                # not "type 1, unit table, t forms": outside the recipe
                if np.abs(d0[:tab] - 1.0).max() > 1e-14 or np.abs(d0[tab + 1:].imag).max() > 0:
                    return
                ratio = [d0[tab].imag / d0[tab].real] + [float(x.real) for x in d0[tab + 1:]]
                K, SG = [], []
                for t in ratio:
                    k = int(np.argmin(np.abs(np.abs(t) - s[0] / c[0])))
                    K.append(k)
                    SG.append(1.0 if t >= 0 else -1.0)
                grp = self.rot.setdefault((R, tab), {"off": [], "K": [], "SG": []})
                grp["off"].append(offb)
                grp["K"].append(K)
                grp["SG"].append(SG)
                for b in range(self.batch):
                    self.const[offb + b * ln: offb + b * ln + tab] = 1.0
            else:
                if np.abs(d0.imag).max() > 0:
                    return
                idx = np.array([int(np.argmin(np.abs(np.abs(v) - tab0))) for v in d0.real])
                sg = np.where(d0.real < 0, -1.0, 1.0)
                self.plain.append((offb, ln, idx, sg))
        for grp in self.rot.values():
            grp["off"] = np.asarray(grp["off"], dtype=np.int64)
            grp["K"] = np.asarray(grp["K"], dtype=np.int64)
            grp["SG"] = np.asarray(grp["SG"], dtype=np.float64)
        # the recipe must reproduce the probe's SECOND member
        self.batch, keep = 2, (self.batch, self.total, self.const, [(o, ln, i, g) for o, ln, i, g in self.plain],
                               {k: dict(v) for k, v in self.rot.items()})
        # (offsets of the two-member program for the check)
        self.total = int(prog.mats.size)
        self.const = np.zeros(self.total, dtype=np.complex128)
        remap = {int(gb[gi]["mat_off"]): int(g2[gi]["mat_off"]) for gi in range(len(g2))}
        self.plain = [(remap[o], ln, i, g) for o, ln, i, g in keep[3]]
        for key, v in keep[4].items():
            self.rot[key] = dict(v, off=np.asarray([remap[int(o)] for o in v["off"]], dtype=np.int64))
            for o in self.rot[key]["off"]:
                for b in range(2):
                    self.const[int(o) + b * (key[1] + key[0]): int(o) + b * (key[1] + key[0]) + key[1]] = 1.0
        for gi in range(len(g2)):
            if int(g2[gi]["mat_bstride"]) == 0:
                o2 = int(g2[gi]["mat_off"])
                nxt = int(g2[gi + 1]["mat_off"]) if gi + 1 < len(g2) else int(prog.mats.size)
                self.const[o2:nxt] = prog.mats[o2:nxt]
        self.ok = True
        got = self.mats(probe)
        good = got is not None and np.abs(got - prog.mats).max() < 1e-13
        self.batch, self.total, self.const, self.plain, self.rot = keep
        self.ok = bool(good)

    def mats(self, params: np.ndarray, out: Optional[np.ndarray] = None) -> Optional[np.ndarray]:
        """The matrix buffer for ``params`` [B, (layers+1)*n].  ``out``: a buffer (complex64 or complex128, e.g. the pinned
        staging buffer of the resident program) that already holds the parameter-independent entries from an earlier call
        -- only the parameter-dependent ones are rewritten."""
        if not self.ok:
            return None
        p = np.asarray(params, dtype=np.float64).reshape(self.batch, -1)
        B = self.batch
        c, s = np.cos(0.5 * p), np.sin(0.5 * p)
        for (R, tab), grp in self.rot.items():       # (before anything is written: the call may have to be refused)
            if R > 1:
                a, r = c[:, grp["K"][:, 1:]], s[:, grp["K"][:, 1:]]
                if not np.all(np.abs(a) * self.t_max >= np.abs(r)):
                    return None                      # a layer needs the c form: lower this call generically
        if out is None:
            out = self.const.copy()
        tabv = np.concatenate([np.zeros((B, 1)), np.ones((B, 1)), c, s], axis=1)
        for off, ln, idx, sg in self.plain:
            out[off:off + B * ln].reshape(B, ln)[:, :] = tabv[:, idx] * sg
        for (R, tab), grp in self.rot.items():
            a = c[:, grp["K"]]                       # [B, G, R]
            r = s[:, grp["K"]] * grp["SG"]
            with np.errstate(all="ignore"):
                t = r[:, :, 1:] / a[:, :, 1:]
            scale = np.prod(a[:, :, 1:], axis=2)
            c0 = (a[:, :, 0] + 1j * r[:, :, 0]) * scale
            ln = tab + R
            for g, off in enumerate(grp["off"]):
                blk = out[off:off + B * ln].reshape(B, ln)
                blk[:, tab] = c0[:, g]
                blk[:, tab + 1:] = t[:, g, :]
        return out


class BatchedAnsatz:
    def __init__(self, n: int, layers: int, batch: int, *, device: str | torch.device = "cuda",
                 dtype: torch.dtype = torch.complex64, tile: Optional[TileConfig] = None) -> None:
        self.n, self.layers, self.batch = int(n), int(layers), int(batch)
        self.device = torch.device(device)
        self.dtype = dtype
        self.itemsize = 16 if dtype == torch.complex128 else 8
        self.tile = tile or default_tile(self.n, self.itemsize, self.batch)
        _lib.ensure_device(self.device.index or 0)
        self.state = torch.empty((self.batch, 1 << self.n), dtype=dtype, device=self.device)
        self.h2d_bytes = 0
        self.passes = 0
        self._refill: Optional[HweRyRefill] = None
        self._prog = self._dp = None
        self._host_view = None
        self.fast_calls = 0

    def run(self, params: np.ndarray) -> torch.Tensor:
        """|psi_b> = ansatz(params[b]) |0..0> for every batch member, in place on the device.  The first call lowers the
        circuit through the generic pipeline and derives the refill recipe (HweRyRefill); later calls only rewrite the
        matrix buffer of the resident program."""
        mats = None
        if self._refill is not None and self._refill.ok:
            if self._host_view is None:   # first refill: the whole buffer, parameter-independent entries included
                mats = self._refill.mats(params)
                if mats is not None:
                    self._dp.fill_host(mats)
                    self._host_view = self._dp.mats_host.numpy()[: mats.size]
            else:                         # later: only the parameter-dependent entries, straight into the pinned buffer
                mats = self._refill.mats(params, out=self._host_view)
        if mats is not None:
            prog, dp = self._prog, self._dp
            dp.upload()
            self.fast_calls += 1
        else:
            gates = fuse(hwe_ry_gates(self.n, self.layers, np.asarray(params)))
            prog = compile_program(gates, self.n, self.tile, batch_mats=self.batch, itemsize=self.itemsize)
            dp = P.DeviceProgram(prog, self.device, self.dtype)
            if self._refill is None and self.batch > 1:
                self._refill = HweRyRefill(self.n, self.layers, prog, self.batch, self.tile, self.itemsize)
                if self._refill.ok:
                    self._prog, self._dp = prog, dp
        ptr, n, b, dt, stream = P._prep(self.state)
        _lib.check(_lib.load().tqb_init_basis(ptr, n, b, dt, 0, 0, stream))
        dp.run(self.state)
        self.h2d_bytes = dp.h2d_bytes
        self.passes = prog.n_passes
        return self.state

    def expvals(self, ham: PauliSum) -> torch.Tensor:
        return ham.expectation(self.state).real

    def sample(self, uniforms: torch.Tensor) -> torch.Tensor:
        return P.sample(self.state, uniforms)
