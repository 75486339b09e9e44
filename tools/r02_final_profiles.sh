#!/bin/bash
# launch list of the bench command + ncu --set full captures of the dominant kernels, final library
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --layers 20 --no-cpu --no-extras > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tqb_spec_pass -c 6 -f -o gpurun_out/r02_spec_final python tools/pass_times.py hea30 6 6 ncu > /dev/null 2>&1
echo "ncu hea30 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tqb_spec_pass -c 1 -f -o gpurun_out/r02_trotter_pass4_final python tools/pass_times.py trotter 1 4 ncu > /dev/null 2>&1
echo "ncu trotter rc=$?"
ls -la gpurun_out/r02_launches.csv gpurun_out/r02_spec_final.ncu-rep gpurun_out/r02_trotter_pass4_final.ncu-rep
