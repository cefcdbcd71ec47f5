#!/bin/bash
# one-GPU session: GPU test tier, leap bench (kernel-only line + phase timers), default bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r02_pytest_gpu.log 2>&1
cat gpurun_out/r02_pytest_gpu.log
( timeout 300 python bench.py --workload leap_cube_mppi --steps 10 --warmup 3 --no-extras ) > gpurun_out/r02_bench_leap_hh.json 2> gpurun_out/r02_bench_leap_hh.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_leap_hh.json').read().strip().splitlines()[-1])
print('leap ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'overflows', d.get('contact_overflows'))
PY
( timeout 600 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -3 gpurun_out/r02_bench_default.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default: value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'p50', d['e2e'].get('plan_latency_p50_ms'))
for k, v in (d.get('also') or {}).items():
    print('also', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'contact_overflows', 'error')}, (v.get('e2e') or {}).get('plan_latency_p50_ms'))
print('cpu_baseline', d.get('cpu_baseline'))
PY
