#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_fr3.py -q 2>&1 | tail -3 ) > gpurun_out/pytest_fr3.log 2>&1
( timeout 100 python bench.py --workload fr3_pick_cem --steps 10 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3.json 2> gpurun_out/bench_fr3.err
( timeout 100 python bench.py --workload fr3_pick_cem_grasp --steps 10 --warmup 3 --cpu-budget 4 ) > gpurun_out/bench_fr3_grasp.json 2> gpurun_out/bench_fr3_grasp.err
cat gpurun_out/pytest_fr3.log
python - <<'PY'
import json
for n in ('fr3', 'fr3_grasp'):
    try:
        d = json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round((d.get('cpu_baseline') or {}).get('value', 0)))
    except Exception as e:
        print(n, 'failed', e, open(f'gpurun_out/bench_{n}.err').read()[-300:])
PY
