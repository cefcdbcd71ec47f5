#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python -m pytest tests -m gpu -q 2>&1 | tail -3 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 100 python bench.py --steps 100 --warmup 10 --cpu-budget 1 ) > gpurun_out/bench_cartpole.json 2> gpurun_out/bench_cartpole.err
( timeout 100 python bench.py --workload leap_cube_mppi --steps 10 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
cat gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for n in ('cartpole', 'leap'):
    try:
        d = json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'overflows', d.get('contact_overflows'))
    except Exception as e:
        print(n, 'failed', e, open(f'gpurun_out/bench_{n}.err').read()[-300:])
PY
