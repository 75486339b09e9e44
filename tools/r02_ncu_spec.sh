#!/bin/bash
# ncu --set full on the specialised pass kernels: c128 (period of 6 passes) and c64
mkdir -p gpurun_out
TQB_JIT=2 timeout 300 python tools/hea_cfg.py 30 12 c128 11:5:128 > gpurun_out/r02_fused_c128.log 2>&1
TQB_JIT=2 timeout 300 python tools/hea_cfg.py 30 12 c64 12:6:128 > gpurun_out/r02_fused_c64.log 2>&1
cat gpurun_out/r02_fused_c128.log gpurun_out/r02_fused_c64.log | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_spec -s 31 -c 6 -o gpurun_out/r02_spec_c128 env TQB_JIT=2 python tools/hea_cfg.py 30 12 c128 11:5:128 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_spec -s 33 -c 3 -o gpurun_out/r02_spec_c64 env TQB_JIT=2 python tools/hea_cfg.py 30 12 c64 12:6:128 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
