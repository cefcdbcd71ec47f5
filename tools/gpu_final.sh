#!/bin/bash
# Round-end evidence run: full GPU test tier, bench lines of the four workloads, ncu launch list + full capture of the fr3 kernel,
# compute-sanitizer memcheck of the warp-per-rollout kernels.
mkdir -p gpurun_out
( timeout 500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 200 python bench.py --steps 200 --warmup 20 --cpu-budget 6 ) > gpurun_out/bench_cartpole.json 2> gpurun_out/bench_cartpole.err
( timeout 200 python bench.py --workload cylinder_push_cem --steps 200 --warmup 20 --cpu-budget 6 ) > gpurun_out/bench_cyl.json 2> gpurun_out/bench_cyl.err
( timeout 200 python bench.py --workload leap_cube_mppi --steps 30 --warmup 5 --cpu-budget 6 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
( timeout 200 python bench.py --workload fr3_pick_cem --steps 20 --warmup 5 --cpu-budget 6 ) > gpurun_out/bench_fr3.json 2> gpurun_out/bench_fr3.err
( timeout 100 python bench.py --workload fr3_pick_cem --n-rollouts 64 --steps 20 --warmup 5 --cpu-budget 1 ) > gpurun_out/bench_fr3_n64.json 2> gpurun_out/bench_fr3_n64.err
( timeout 100 python bench.py --workload fr3_pick_cem --n-rollouts 4096 --steps 8 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_n4096.json 2> gpurun_out/bench_fr3_n4096.err
( B200MPC_FR3_PROF=1 timeout 100 python bench.py --workload fr3_pick_cem --steps 3 --warmup 1 --cpu-budget 1 ) > gpurun_out/prof_fr3.json 2> gpurun_out/prof_fr3.err
( timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fr3.csv python bench.py --workload fr3_pick_cem --steps 2 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_launch.log 2>&1
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:fr3_rollout_kernel -c 1 -f -o gpurun_out/fr3_full3 python bench.py --workload fr3_pick_cem --steps 1 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_full.log 2>&1
( timeout 100 ncu -i gpurun_out/fr3_full3.ncu-rep --page raw --csv ) > gpurun_out/fr3_full3_raw.csv 2>&1
( timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fr3.py -q -k "rollout_matches_oracle or plan_costs" 2>&1 | tail -15 ) > gpurun_out/sanitizer_fr3.log 2>&1
( timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -k "leap_rollout_matches_oracle" 2>&1 | tail -15 ) > gpurun_out/sanitizer_leap.log 2>&1
cat gpurun_out/pytest_gpu.log
grep fr3_prof gpurun_out/prof_fr3.err
tail -n 4 gpurun_out/sanitizer_fr3.log gpurun_out/sanitizer_leap.log
python - <<'PY'
import json
for n in ('cartpole', 'cyl', 'leap', 'fr3', 'fr3_n64', 'fr3_n4096'):
    try:
        d = json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), 'e2e', round(d['e2e']['value']), 'cpu', round((d.get('cpu_baseline') or {}).get('value', 0)))
    except Exception as e:
        print(n, 'failed', e)
PY
