#!/bin/bash
mkdir -p gpurun_out
( timeout 400 python -m pytest tests -m gpu -q 2>&1 | tail -6 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 100 python bench.py --workload leap_cube_mppi --steps 30 --warmup 5 --cpu-budget 2 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
( timeout 100 python bench.py --workload fr3_pick_cem --steps 20 --warmup 5 --cpu-budget 2 ) > gpurun_out/bench_fr3.json 2> gpurun_out/bench_fr3.err
cat gpurun_out/pytest_gpu.log
python - <<'PY'
import json
for n in ('leap', 'fr3'):
    try:
        d = json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(n, 'failed', e)
PY
