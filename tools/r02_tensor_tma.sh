#!/bin/bash
mkdir -p gpurun_out
LOG=gpurun_out/r02_tensor_tma.log
: > $LOG
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -m gpu -k "specialised or dropin" >> $LOG 2>&1
echo "tests rc=$?" >> $LOG
run() { label=$1; shift; echo "== $label" >> $LOG; timeout -s ABRT 200 env "$@" >> $LOG 2>&1; echo "rc=$?" >> $LOG; }
run "tensor c128" TQB_JIT=2 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "bulk   c128" TQB_JIT=2 TQB_TENSOR_TMA=0 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "tensor c128 transfer-only" TQB_JIT=2 TQB_DBG=1 python -X faulthandler tools/hea_cfg.py 30 12 c128 11:5:128
run "tensor c64" TQB_JIT=2 python -X faulthandler tools/hea_cfg.py 30 12 c64 12:6:128
run "bulk   c64" TQB_JIT=2 TQB_TENSOR_TMA=0 python -X faulthandler tools/hea_cfg.py 30 12 c64 12:6:128
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/r02_tensor_launches.csv env TQB_JIT=2 python tools/hea_cfg.py 30 12 c128 11:5:128 > /dev/null 2>&1
grep -v "^$" $LOG | tail -30
