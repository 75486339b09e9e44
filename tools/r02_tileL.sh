#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r02_tileL.log; : > $LOG
run() { label=$1; shift; echo "== $label" >> $LOG; timeout -s ABRT 400 env "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run "c128 L=5,4,3 (20 layers)" TQB_JIT=2 python -X faulthandler tools/hea_cfg.py 30 20 c128 11:5:128 11:4:128 11:3:128
run "c64 L=6,5,4 (20 layers)" TQB_JIT=2 python -X faulthandler tools/hea_cfg.py 30 20 c64 12:6:128 12:5:128 12:4:128
grep -v "^$" $LOG | tail -12
