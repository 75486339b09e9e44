#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/grad_scale.py 30 20 > gpurun_out/r02_grad_scale30.log 2>&1; tail -4 gpurun_out/r02_grad_scale30.log
