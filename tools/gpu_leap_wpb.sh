#!/bin/bash
mkdir -p gpurun_out
for w in 7 8 4 2 1; do
  ( B200MPC_LEAP_WPB=$w timeout 100 python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_leap_wpb$w.json 2> gpurun_out/bench_leap_wpb$w.err
done
( timeout 200 python -m pytest tests/test_gpu_parity.py -q -k leap 2>&1 | tail -3 ) > gpurun_out/pytest_leap.log 2>&1
cat gpurun_out/pytest_leap.log
python - <<'PY'
import json
for w in (7, 8, 4, 2, 1):
    try:
        d = json.load(open(f'gpurun_out/bench_leap_wpb{w}.json'))
        print('wpb', w, 'ms/step', round(d['ms_per_step'], 3), 'rollouts/s', round(d['value']))
    except Exception as e:
        print(w, 'failed', e, open(f'gpurun_out/bench_leap_wpb{w}.err').read()[-300:])
PY
