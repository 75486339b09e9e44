"""tyxonq_b200 -- B200-native (sm_100a) statevector engine behind TyxonQ's device/numerics API.

Scope: the one data-parallel hot path of QureGenAI-Biotech/TyxonQ -- gate application on the
2^n-amplitude state, Pauli-sum expectation, shot sampling and VQE parameter gradients.  Python
host code (this package) drives hand-written CUDA kernels through the C ABI declared in
``include/tyxonq_b200.h``; there is no CPU fallback.

Public surface (mirrors the reference's names):
  StatevectorEngine           devices/simulators/statevector/engine.py
  kernels.*                   libs/quantum_library/kernels/statevector.py
  PauliSum                    libs/quantum_library/kernels/pauli.py + dynamics.expectation
  DensityMatrixEngine         devices/simulators/density_matrix/engine.py (rho as a 2n-bit vector on the same kernels)
  measure.GroupedMeasurement  shots > 0 energies: hamiltonian_grouping.py + counts_expval.py + the device runtimes' loops
  noise.TrajectoryBatch       batched Monte-Carlo Kraus trajectories (kernels/statevector.py:132-218)
  B200Backend                 numerics/api.py ArrayBackend, for ``tq.set_backend(B200Backend())``
  install()                   route ``device="statevector"`` of a live TyxonQ install to this engine
"""
from __future__ import annotations

from ._lib import TqbError, launch_count, load  # noqa: F401

__version__ = "0.1.0"

_LAZY = {
    "StatevectorEngine": ("engine", "StatevectorEngine"),
    "PauliSum": ("pauli", "PauliSum"),
    "B200Backend": ("backend", "B200Backend"),
    "DensityMatrixEngine": ("density", "DensityMatrixEngine"),
    "ShardedStatevectorEngine": ("sharded_engine", "ShardedStatevectorEngine"),
    "install": ("install", "install"),
    "uninstall": ("install", "uninstall"),
}


def __getattr__(name: str):
    if name in _LAZY:
        import importlib
        mod, attr = _LAZY[name]
        obj = getattr(importlib.import_module(f"{__name__}.{mod}"), attr)
        # importing the submodule ``install`` binds the MODULE to this package's attribute ``install``: rebind the
        # function, or a second ``tyxonq_b200.install()`` would find the module
        globals()[name] = obj
        if mod == "install":
            globals()["install"] = getattr(importlib.import_module(f"{__name__}.install"), "install")
        return obj
    if name in ("kernels", "engine", "pauli", "program", "planner", "gates", "autograd", "ucc", "vqe", "sharded", "sharded_engine", "circuits", "backend", "batched", "measure", "noise", "density"):
        import importlib
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
