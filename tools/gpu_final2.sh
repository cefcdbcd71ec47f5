#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -6 ) > gpurun_out/smoke.log 2>&1
( timeout 200 python bench.py --steps 200 --warmup 20 --cpu-budget 6 ) > gpurun_out/bench_cartpole.json 2> gpurun_out/bench_cartpole.err
( timeout 200 python bench.py --workload cylinder_push_cem --steps 200 --warmup 20 --cpu-budget 6 ) > gpurun_out/bench_cyl.json 2> gpurun_out/bench_cyl.err
( timeout 200 python bench.py --workload leap_cube_mppi --steps 30 --warmup 5 --cpu-budget 6 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
( timeout 200 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/bench_reference_cartpole.json 2> gpurun_out/bench_reference_cartpole.err
( timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_leap.csv python bench.py --workload leap_cube_mppi --steps 2 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_launch_leap.log 2>&1
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:leap_rollout_kernel -c 1 -f -o gpurun_out/leap_full4 python bench.py --workload leap_cube_mppi --steps 1 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_leap.log 2>&1
( timeout 100 ncu -i gpurun_out/leap_full4.ncu-rep --page raw --csv ) > gpurun_out/leap_full4_raw.csv 2>&1
( B200MPC_LEAP_PROF=1 timeout 100 python bench.py --workload leap_cube_mppi --steps 3 --warmup 1 --cpu-budget 1 ) > gpurun_out/prof_leap.json 2> gpurun_out/prof_leap.err
cat gpurun_out/smoke.log
grep leap_prof gpurun_out/prof_leap.err
python - <<'PY'
import json
for n in ('cartpole', 'cyl', 'leap', 'reference_cartpole'):
    try:
        d = json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round((d.get('cpu_baseline') or {}).get('value', 0)))
    except Exception as e:
        print(n, 'failed', e)
PY
