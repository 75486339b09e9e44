"""Route a live TyxonQ install to the B200 engine (seams B1/B2 of SURVEY.md section 8b).

``install()`` rebinds, inside an importable ``tyxonq`` package:
  * ``devices.simulators.driver._select_engine``  -> returns our ``StatevectorEngine`` for
    "statevector" (reference driver.py:20-30, 96-97: the engine is constructed with no args)
  * ``devices.simulators.statevector.engine.StatevectorEngine`` -> our class (used by
    ``Circuit.state``, core/ir/circuit.py:492-494)
  * the kernel functions of ``libs.quantum_library.kernels.statevector`` (looked up lazily by
    ``Circuit._expectation_statevector`` and the chem numerics)
  * optionally (seam B3) the process-global numerics backend -> ``B200Backend``
TyxonQ itself is not a dependency of this package; ``install()`` raises ImportError without it.
"""
from __future__ import annotations

from typing import Any, Dict

_saved: Dict[str, Any] = {}


def _multi_rank() -> bool:
    import torch.distributed as dist
    return bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)


def install(set_backend: bool = False, density: bool = False, sharded: bool = False) -> None:
    """``set_backend=True`` additionally makes ``B200Backend`` the process-global numerics backend (seam B3,
    numerics/__init__.py:20-36), so that ``Circuit.state()`` returns device tensors and ``K.value_and_grad`` runs the
    adjoint sweep.  ``density=True`` also routes ``device="density_matrix"`` (driver.py:20-30) to the
    density-matrix engine that rides on the same kernels (density.py; at most 17 qubits).  ``sharded=True``: when the process
    runs under ``torchrun`` with an initialised process group of more than one rank, "statevector" resolves to
    ``ShardedStatevectorEngine`` (sharded_engine.py): every rank runs the same script, ``run`` is collective."""
    import importlib
    drv = importlib.import_module("tyxonq.devices.simulators.driver")
    eng_mod = importlib.import_module("tyxonq.devices.simulators.statevector.engine")
    ker_mod = importlib.import_module("tyxonq.libs.quantum_library.kernels.statevector")
    from . import kernels as K
    from .engine import StatevectorEngine

    if _saved:
        return
    _saved["select"] = drv._select_engine
    _saved["engine"] = eng_mod.StatevectorEngine
    _saved["kernels"] = {k: getattr(ker_mod, k) for k in K.__all__}

    ref_select = drv._select_engine

    def _select_engine(device: str):
        name = device.split("::")[-1] if "::" in device else device
        if name in ("simulator:statevector", "statevector"):
            if sharded and _multi_rank():
                from .sharded_engine import ShardedStatevectorEngine
                return ShardedStatevectorEngine
            return StatevectorEngine
        if density and name in ("simulator:density_matrix", "density_matrix"):
            from .density import DensityMatrixEngine
            return DensityMatrixEngine
        return ref_select(device)

    drv._select_engine = _select_engine
    eng_mod.StatevectorEngine = StatevectorEngine
    for k in K.__all__:
        setattr(ker_mod, k, getattr(K, k))
    if set_backend:
        from .backend import B200Backend
        num = importlib.import_module("tyxonq.numerics")
        num.set_backend(B200Backend())
        _saved["backend"] = True


def uninstall() -> None:
    if not _saved:
        return
    import importlib
    drv = importlib.import_module("tyxonq.devices.simulators.driver")
    eng_mod = importlib.import_module("tyxonq.devices.simulators.statevector.engine")
    ker_mod = importlib.import_module("tyxonq.libs.quantum_library.kernels.statevector")
    drv._select_engine = _saved["select"]
    eng_mod.StatevectorEngine = _saved["engine"]
    for k, v in _saved["kernels"].items():
        setattr(ker_mod, k, v)
    if _saved.get("backend"):
        importlib.import_module("tyxonq.numerics").set_backend("numpy")
    _saved.clear()
