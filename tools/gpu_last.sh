mkdir -p gpurun_out
( timeout 90 python bench.py --steps 100 --warmup 10 --cpu-budget 1 ) > gpurun_out/bench_cartpole.json 2> gpurun_out/bench_cartpole.err
( timeout 60 python bench.py --workload leap_cube_mppi --steps 10 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
( timeout 60 python bench.py --workload fr3_pick_cem --steps 5 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3.json 2> gpurun_out/bench_fr3.err
python - <<'PY'
import json
for n in ('cartpole', 'leap', 'fr3'):
    try:
        d = json.load(open(f'gpurun_out/bench_{n}.json'))
        print(n, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), 'plan_latency', d['plan_latency_c1'].get('workload'))
    except Exception as e:
        print(n, 'failed', e, open(f'gpurun_out/bench_{n}.err').read()[-300:])
PY
