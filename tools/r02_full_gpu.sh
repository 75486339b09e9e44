#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_full.log; tail -6 gpurun_out/r02_gpu_tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
