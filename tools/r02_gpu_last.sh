#!/bin/bash
# last, short one-GPU pass: GPU test tier, smoke, default bench line, Controller latency breakdown
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 ) > gpurun_out/r02_pytest_gpu.log 2>&1; cat gpurun_out/r02_pytest_gpu.log
( timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 ) > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
( timeout 400 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
python tools/r02_latency.py > gpurun_out/r02_latency.txt 2>&1
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default: value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'p50', d['e2e'].get('plan_latency_p50_ms'), 'c1 p50', d.get('plan_latency_c1_p50_ms'))
for k, v in (d.get('also') or {}).items():
    print('also', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'contact_overflows', 'error')}, (v.get('e2e') or {}).get('plan_latency_p50_ms'))
PY
grep "fast=True" gpurun_out/r02_latency.txt
