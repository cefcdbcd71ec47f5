#!/bin/bash
# fr3 optimisation loop: parity, bench at three sizes, sync-mode sweep, phase profile (clock64 timers), full GPU suite
mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_fr3.py -q 2>&1 | tail -30 ) > gpurun_out/pytest_fr3.log 2>&1
for n in 1024 64 4096; do
  ( timeout 150 python bench.py --workload fr3_pick_cem --n-rollouts $n --steps 8 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_n$n.json 2> gpurun_out/bench_fr3_n$n.err
done
for sm in 0 1 2; do
  ( B200MPC_FR3_SYNC=$sm timeout 150 python bench.py --workload fr3_pick_cem --steps 8 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_sync$sm.json 2> gpurun_out/bench_fr3_sync$sm.err
done
( B200MPC_FR3_PROF=1 timeout 150 python bench.py --workload fr3_pick_cem --steps 3 --warmup 1 --cpu-budget 1 ) > gpurun_out/prof_fr3.json 2> gpurun_out/prof_fr3.err
( timeout 150 python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
( timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fr3.py 2>&1 | tail -30 ) > gpurun_out/pytest_gpu.log 2>&1
cat gpurun_out/pytest_fr3.log gpurun_out/pytest_gpu.log
grep fr3_prof gpurun_out/prof_fr3.err
python - <<'PY'
import json
for n in ('n64', 'n1024', 'n4096', 'sync0', 'sync1', 'sync2'):
    try:
        d = json.load(open(f'gpurun_out/bench_fr3_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 3), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(n, 'failed', e)
d = json.load(open('gpurun_out/bench_leap.json'))
print('leap ms/step', round(d['ms_per_step'], 3), 'rollouts/s', round(d['value']))
PY
