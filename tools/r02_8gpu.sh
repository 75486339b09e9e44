#!/bin/bash
# 8 GPUs: bench line (HEA-33 complex128 weak scaling + config 4: 36-qubit complex64 Trotter)
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench$N.json 2> gpurun_out/r02_bench$N.err; echo "bench rc=$?"; grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02_bench$N.err | tail -5
python - $N <<'PY'
import json,sys
N=sys.argv[1]
try:
    d=json.loads([l for l in open(f'gpurun_out/r02_bench{N}.json').read().strip().splitlines() if l.startswith('{')][-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'passes', d['config'].get('passes'), 'frac', d['roofline']['frac'])
    print('nvlink', d.get('nvlink'))
    print('e2e', d['e2e']['value'])
    print('extras', json.dumps(d.get('extras'))[:1500])
except Exception as e: print('parse failed', e)
PY
