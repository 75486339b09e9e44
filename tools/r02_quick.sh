#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_measure.py tests/test_gpu_grad.py -x -q -m gpu -k "pauli or sampl or shift or ucc or tfim or counts or batched" > gpurun_out/r02_quick_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_quick_tests.log; tail -4 gpurun_out/r02_quick_tests.log
timeout 300 python tools/pauli_cfg5.py > gpurun_out/r02_cfg5.log 2>&1; cat gpurun_out/r02_cfg5.log | tail -8
