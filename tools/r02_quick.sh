#!/bin/bash
# parity tests of the pass kernels + the default bench (every step bounded)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/r02_parity_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_parity_tests.log
timeout 250 python tools/pass_times.py trotter 15 > gpurun_out/r02_pass_times_trotter_l0s.log 2>&1; tail -1 gpurun_out/r02_pass_times_trotter_l0s.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','specialised_launches')})
    print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','launch_ms')}, 'compute', d['roofline'].get('compute',{}).get('frac'))
    print('e2e', d['e2e']['value'])
    for k,v in d.get('extras',{}).items(): print(k, json.dumps({a:b for a,b in v.items() if a not in ('cpu_baseline','note','batch')})[:420])
except Exception as e: print('parse failed', e)
PY
