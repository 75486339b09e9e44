"""Route a live TyxonQ install to the B200 engine (seams B1/B2 of SURVEY.md section 8b).

``install()`` rebinds, inside an importable ``tyxonq`` package:
  * ``devices.simulators.driver._select_engine``  -> returns our ``StatevectorEngine`` for
    "statevector" (reference driver.py:20-30, 96-97: the engine is constructed with no args)
  * ``devices.simulators.statevector.engine.StatevectorEngine`` -> our class (used by
    ``Circuit.state``, core/ir/circuit.py:492-494)
  * the kernel functions of ``libs.quantum_library.kernels.statevector`` (looked up lazily by
    ``Circuit._expectation_statevector`` and the chem numerics)
  * ``devices.simulators.driver.run`` -> the same function without the second simulation and the 2^n host copy of
    its ``shots == 0`` epilogue (driver.py:115-126; LazyHostArray from ``lazy_min_qubits`` qubits on)
  * optionally (seam B3) the process-global numerics backend -> ``B200Backend``
TyxonQ itself is not a dependency of this package; ``install()`` raises ImportError without it.
"""
from __future__ import annotations

from typing import Any, Dict

_saved: Dict[str, Any] = {}


def _multi_rank() -> bool:
    import torch.distributed as dist
    return bool(dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1)


def _make_driver_run(drv, ref_run, lazy_min_qubits: int):
    """``devices.simulators.driver.run`` (driver.py:86-142) for the devices this package serves: same result dict,
    but the ``shots == 0`` epilogue (driver.py:115-126) neither simulates a second time (the engine keeps the state of
    ``run``) nor copies 2^n amplitudes to the host for circuits of ``lazy_min_qubits`` or more (LazyHostArray)."""
    from uuid import uuid4

    import numpy as np

    from . import program as P
    from .engine import StatevectorEngine
    from .lazy import LazyHostArray

    def run(device, token=None, *, circuit=None, source=None, shots=1024, **opts):
        Engine = drv._select_engine(device)
        if not (isinstance(Engine, type) and issubclass(Engine, StatevectorEngine)):
            return ref_run(device, token, circuit=circuit, source=source, shots=shots, **opts)
        circuit = drv._qasm_to_ir_if_needed(circuit, source)
        eng = Engine()

        def _one(c):
            error = ""
            out = {}
            exact = not isinstance(shots, (list, tuple)) and int(shots) == 0
            try:
                out = eng.run(c, shots=shots, _keep_state=exact, **opts)
            except Exception as e:   # noqa: BLE001 -- the reference driver reports errors in the result dict
                error = str(e)
            counts = out.get("result") or {}
            expectations = out.get("expectations") or {}
            meta = dict(out.get("metadata", {}))
            prob = None
            statevec = None
            if exact and not error:
                try:
                    psi = eng.state_device(c)
                    n = int(getattr(c, "num_qubits", 0))
                    meta.setdefault("num_qubits", n)
                    if n >= lazy_min_qubits:
                        statevec = LazyHostArray(psi)
                        prob = LazyHostArray(psi, transform=P.probabilities, dtype=np.float64)
                    else:
                        statevec = psi.to(dtype=__import__("torch").complex128).cpu().numpy()
                        prob = np.abs(statevec) ** 2
                except Exception as e:   # noqa: BLE001
                    error = str(e)
                    prob = None
                    statevec = None
            result = {"result": counts, "expectations": expectations, "probabilities": prob, "statevector": statevec,
                      "result_meta": meta, "uni_status": "completed", "error": error}
            return drv.SimTask(id=str(uuid4()), device=device, result=result)

        if isinstance(circuit, (list, tuple)):
            return [_one(c) for c in circuit]
        return [_one(circuit)]

    return run


def install(set_backend: bool = False, density: bool = False, sharded: bool = False, lazy_min_qubits: int | None = None) -> None:
    """``set_backend=True`` additionally makes ``B200Backend`` the process-global numerics backend (seam B3,
    numerics/__init__.py:20-36), so that ``Circuit.state()`` returns device tensors and ``K.value_and_grad`` runs the
    adjoint sweep.  ``density=True`` also routes ``device="density_matrix"`` (driver.py:20-30) to the
    density-matrix engine that rides on the same kernels (density.py; at most 17 qubits).  ``sharded=True``: when the process
    runs under ``torchrun`` with an initialised process group of more than one rank, "statevector" resolves to
    ``ShardedStatevectorEngine`` (sharded_engine.py): every rank runs the same script, ``run`` is collective.
    ``lazy_min_qubits`` (default 26, env TQB_LAZY_MIN_QUBITS): from this size on the ``statevector`` /
    ``probabilities`` entries of a ``shots == 0`` result are LazyHostArray views of the device state (lazy.py).
    Calling ``install`` again with other flags re-routes (the previous routing is undone first)."""
    import importlib
    import os
    drv = importlib.import_module("tyxonq.devices.simulators.driver")
    eng_mod = importlib.import_module("tyxonq.devices.simulators.statevector.engine")
    ker_mod = importlib.import_module("tyxonq.libs.quantum_library.kernels.statevector")
    from . import kernels as K
    from .engine import StatevectorEngine

    flags = (bool(set_backend), bool(density), bool(sharded), lazy_min_qubits)
    if _saved:
        if _saved.get("flags") == flags:
            return
        uninstall()
    if lazy_min_qubits is None:
        lazy_min_qubits = int(os.environ.get("TQB_LAZY_MIN_QUBITS", "26"))
    _saved["flags"] = flags
    _saved["select"] = drv._select_engine
    _saved["run"] = drv.run
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
    drv.run = _make_driver_run(drv, _saved["run"], int(lazy_min_qubits))
    eng_mod.StatevectorEngine = StatevectorEngine
    for k in K.__all__:
        setattr(ker_mod, k, getattr(K, k))
    if set_backend:
        from .backend import B200Backend
        num = importlib.import_module("tyxonq.numerics")
        try:
            _saved["prev_backend"] = num.get_backend(None)
        except Exception:   # noqa: BLE001
            _saved["prev_backend"] = "numpy"
        num.set_backend(B200Backend())
        _saved["backend"] = True


def reference_kernels() -> Dict[str, Any]:
    """The reference's own kernel functions as they were before install() (empty when not installed)."""
    return dict(_saved.get("kernels", {}))


def uninstall() -> None:
    if not _saved:
        return
    import importlib
    drv = importlib.import_module("tyxonq.devices.simulators.driver")
    eng_mod = importlib.import_module("tyxonq.devices.simulators.statevector.engine")
    ker_mod = importlib.import_module("tyxonq.libs.quantum_library.kernels.statevector")
    drv._select_engine = _saved["select"]
    drv.run = _saved["run"]
    eng_mod.StatevectorEngine = _saved["engine"]
    for k, v in _saved["kernels"].items():
        setattr(ker_mod, k, v)
    if _saved.get("backend"):
        importlib.import_module("tyxonq.numerics").set_backend(_saved.get("prev_backend", "numpy"))
    _saved.clear()
