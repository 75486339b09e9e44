#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:expect_pauli_tiled -s 2 -c 2 -f -o gpurun_out/r02_pauli_tiled python tools/pauli_cfg5.py 256 > /dev/null 2>&1
ls -la gpurun_out/r02_pauli_tiled.ncu-rep
