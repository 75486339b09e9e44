#!/bin/bash
# full GPU tier, smoke and the default bench in one call (every step bounded)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_gpu_tests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_gpu_tests_full.log; tail -8 gpurun_out/r02_gpu_tests_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
echo "bench rc=$?"
tail -5 gpurun_out/r02_bench.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/r02_bench.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','specialised_launches')})
    print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','launch_ms')}, 'compute', d['roofline'].get('compute',{}).get('frac'))
    print('e2e', d['e2e'])
    print('cpu', d.get('cpu_baseline'))
    for k,v in d.get('extras',{}).items(): print(k, json.dumps(v)[:900])
except Exception as e: print('parse failed', e)
PY
