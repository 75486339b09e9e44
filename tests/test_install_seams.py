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
        assert ref_engine.StatevectorEngine is ref_cls
        assert ref_kernels.apply_1q_statevector is ref_fn
        assert driver._select_engine("statevector") is ref_cls
    finally:
        sys.path.remove(REF)
