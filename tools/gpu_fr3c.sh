#!/bin/bash
mkdir -p gpurun_out
( timeout 200 python -m pytest tests/test_gpu_fr3.py -q -x 2>&1 | tail -5 ) > gpurun_out/pytest_fr3.log 2>&1
for n in 1024 64 4096; do
  ( timeout 150 python bench.py --workload fr3_pick_cem --n-rollouts $n --steps 8 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_n$n.json 2> gpurun_out/bench_fr3_n$n.err
done
( B200MPC_FR3_PROF=1 timeout 150 python bench.py --workload fr3_pick_cem --steps 3 --warmup 1 --cpu-budget 1 ) > gpurun_out/prof_fr3.json 2> gpurun_out/prof_fr3.err
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:fr3_rollout_kernel -c 1 -f -o gpurun_out/fr3_full2 python bench.py --workload fr3_pick_cem --steps 1 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_full.log 2>&1
cat gpurun_out/pytest_fr3.log
grep fr3_prof gpurun_out/prof_fr3.err
python - <<'PY'
import json
for n in ('n64', 'n1024', 'n4096'):
    try:
        d = json.load(open(f'gpurun_out/bench_fr3_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 3), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']))
    except Exception as e:
        print(n, 'failed', e)
PY
