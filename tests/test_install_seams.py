"""install()/uninstall() rebind the reference's three seams (only runs where TyxonQ is importable:
the build container mounts it at /root/reference; the GPU box does not have it)."""
from __future__ import annotations

import os
import sys

import pytest

REF = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_install_rebinds_and_restores():
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        import tyxonq  # noqa: F401
        from tyxonq.devices.simulators import driver
        from tyxonq.devices.simulators.statevector import engine as ref_engine
        from tyxonq.libs.quantum_library.kernels import statevector as ref_kernels
        import tyxonq_b200
        from tyxonq_b200 import kernels as K
        ref_cls = ref_engine.StatevectorEngine
        ref_fn = ref_kernels.apply_1q_statevector
        tyxonq_b200.install()
        try:
            assert driver._select_engine("statevector") is tyxonq_b200.StatevectorEngine
            assert driver._select_engine("simulator::statevector") is tyxonq_b200.StatevectorEngine
            assert ref_engine.StatevectorEngine is tyxonq_b200.StatevectorEngine
            assert ref_kernels.apply_1q_statevector is K.apply_1q_statevector
            assert driver._select_engine("density_matrix").__name__ == "DensityMatrixEngine"
            # the class contract the driver relies on (driver.py:96-97, engine.py:35-41)
            eng_cls = driver._select_engine("statevector")
            assert eng_cls.name == "statevector" and eng_cls.capabilities == {"supports_shots": True}
            for meth in ("run", "state", "probability", "amplitude", "perfect_sampling", "expval"):
                assert callable(getattr(eng_cls, meth))
        finally:
            tyxonq_b200.uninstall()
        # sharded=True: the sharded engine only when a process group with more than one rank is live
        inst = sys.modules["tyxonq_b200.install"]
        tyxonq_b200.install(sharded=True)
        try:
            assert driver._select_engine("statevector") is tyxonq_b200.StatevectorEngine
            real = inst._multi_rank
            inst._multi_rank = lambda: True
            try:
                cls = driver._select_engine("statevector")
                assert cls is tyxonq_b200.ShardedStatevectorEngine and cls.name == "statevector"
                for meth in ("run", "state", "amplitude", "expval", "state_shard"):
                    assert callable(getattr(cls, meth))
            finally:
                inst._multi_rank = real
        finally:
            tyxonq_b200.uninstall()
        assert ref_engine.StatevectorEngine is ref_cls
        assert ref_kernels.apply_1q_statevector is ref_fn
        assert driver._select_engine("statevector") is ref_cls
    finally:
        sys.path.remove(REF)


def test_backend_declares_the_whole_array_protocol():
    """numerics/api.py:19-121: every attribute and method of ArrayBackend exists on B200Backend (class level: no GPU
    needed); when the reference is mounted the list is checked against the live Protocol as well."""
    from tyxonq_b200.backend import PROTOCOL_ATTRS, PROTOCOL_METHODS, B200Backend
    for a in PROTOCOL_ATTRS:
        assert hasattr(B200Backend, a), a
    for m in PROTOCOL_METHODS + ["value_and_grad", "vmap", "jit", "set_dtype", "dot", "copy", "allclose", "isclose"]:
        assert callable(getattr(B200Backend, m)), m
    assert B200Backend.name == "b200"
    if os.path.isdir(REF):
        sys.path.insert(0, REF)
        try:
            from tyxonq.numerics.api import ArrayBackend
            declared = [k for k, v in vars(ArrayBackend).items() if callable(v) and not k.startswith("_")]
            assert declared and set(declared) <= set(PROTOCOL_METHODS), set(declared) - set(PROTOCOL_METHODS)
        finally:
            sys.path.remove(REF)


def test_backend_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from tyxonq_b200 import B200Backend, TqbError
    with pytest.raises(TqbError):
        B200Backend()


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference not mounted")
def test_engine_follows_the_global_backend_of_a_live_install():
    """StatevectorEngine(backend_name=None) resolves through tq's process-global backend (api.py:230-234)."""
    sys.path.insert(0, REF)
    sys.dont_write_bytecode = True
    try:
        import tyxonq as tq
        from tyxonq_b200.engine import StatevectorEngine

        class _Fake:
            name = "b200"
        try:
            tq.set_backend("pytorch")
            assert StatevectorEngine().backend_name == "pytorch"
            tq.set_backend(_Fake())
            assert StatevectorEngine().backend_name == "b200"
            assert StatevectorEngine(backend_name="numpy").backend_name == "numpy"
        finally:
            tq.set_backend("numpy")
        assert StatevectorEngine().backend_name == "numpy"
    finally:
        sys.path.remove(REF)
