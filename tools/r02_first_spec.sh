#!/bin/bash
# round 2: specialised kernels -- HEA-30 throughput generic vs specialised, ablation (every step bounded)
mkdir -p gpurun_out
LOG=gpurun_out/r02_hea_first.log
: > $LOG
run() {  # run <label> <env...> -- args
  label=$1; shift
  echo "== $label" >> $LOG
  timeout -s ABRT 200 env "$@" >> $LOG 2>&1
  echo "rc=$?" >> $LOG
}
run "JIT=0 c128" TQB_JIT=0 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "JIT=2 c128" TQB_JIT=2 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "JIT=2 transfer-only" TQB_JIT=2 TQB_DBG=1 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "JIT=2 compute-only" TQB_JIT=2 TQB_DBG=6 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "JIT=2 c64" TQB_JIT=2 python -X faulthandler tools/hea_cfg.py 30 12 c64 12:6:128
run "JIT=0 c64" TQB_JIT=0 python -X faulthandler tools/hea_cfg.py 30 12 c64 12:6:128
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 80 --csv --log-file gpurun_out/r02_first_launches.csv env TQB_JIT=2 python tools/hea_cfg.py 30 12 c128 11:5:128 > /dev/null 2>&1
cat $LOG | grep -v "^$" | tail -40
