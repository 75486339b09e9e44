#!/bin/bash
# 2 GPUs: sharded == unsharded check, reference arm under torchrun, bench line with config 4 extra
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/sharded_check.py 26 4 > gpurun_out/r02_sharded_check2.log 2>&1; echo "check rc=$?"; tail -4 gpurun_out/r02_sharded_check2.log
echo skip ref
timeout 900 $TR bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench rc=$?"; tail -3 gpurun_out/r02_bench2.err
python - <<'PY'
import json
try:
    d=json.loads([l for l in open('gpurun_out/r02_bench2.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config'].get('passes'), d['roofline']['frac'])
    print('nvlink', d.get('nvlink'))
    print('e2e', d['e2e'])
    print('extras', json.dumps(d.get('extras'))[:1500])
except Exception as e: print('parse failed', e)
PY
