#!/bin/bash
mkdir -p gpurun_out
timeout 250 python tools/pass_times.py trotter 15 > gpurun_out/r02_pass_times_trotter_looped.log 2>&1; tail -1 gpurun_out/r02_pass_times_trotter_looped.log
timeout 200 python tools/pass_times.py hea30c64 14 > gpurun_out/r02_pass_times_hea30c64_looped.log 2>&1; tail -1 gpurun_out/r02_pass_times_hea30c64_looped.log
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_looped_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_looped_tests.log
