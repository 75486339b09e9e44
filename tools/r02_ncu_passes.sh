#!/bin/bash
# ncu --set full of one specialised pass: Trotter-33 c64 pass 4 (15 four-layer chains), looped code shape
mkdir -p gpurun_out
timeout 280 ncu --set full --clock-control none --import-source on -k regex:tqb_spec_pass -c 1 -f -o gpurun_out/r02_trotter_pass4_looped python tools/pass_times.py trotter 1 4 ncu > gpurun_out/r02_ncu_trotter.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
