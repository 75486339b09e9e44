#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_grad.py tests/test_gpu_parity.py -x -q -m gpu -k "resident or tfim or value_and_grad or replicas" > gpurun_out/r02_vqe_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r02_vqe_tests.log; tail -5 gpurun_out/r02_vqe_tests.log
timeout 300 python tools/vqe_evals.py > gpurun_out/r02_vqe_evals.log 2>&1; cat gpurun_out/r02_vqe_evals.log | tail -8
