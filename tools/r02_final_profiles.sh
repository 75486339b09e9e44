#!/bin/bash
# launch list of the bench command + one ncu --set full capture of the dominant kernel (period of 6 passes), final library
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --layers 20 --no-cpu --no-extras > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tqb_spec -s 31 -c 6 -o gpurun_out/r02_spec_final env TQB_JIT=2 python tools/hea_cfg.py 30 12 c128 11:5:128 > /dev/null 2>&1
echo "ncu full rc=$?"
ls -la gpurun_out/r02_launches.csv gpurun_out/r02_spec_final.ncu-rep
