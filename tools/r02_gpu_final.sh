#!/bin/bash
# final one-GPU pass of round 2: GPU test tier, smoke, default bench line (C2 + also C3/C4), reference arm at every BASELINE config, leap phase timers
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r02_pytest_gpu.log 2>&1
cat gpurun_out/r02_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 ) > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
( timeout 600 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -2 gpurun_out/r02_bench_default.err
for w in cartpole_ps cartpole_mppi cylinder_push_cem leap_cube_mppi; do
  ( timeout 300 python bench.py --impl reference --workload $w --steps 30 --warmup 3 ) > gpurun_out/r02_bench_reference_$w.json 2> gpurun_out/r02_bench_reference_$w.err
done
( timeout 300 python bench.py --workload cartpole_ps --steps 100 --warmup 10 --no-extras ) > gpurun_out/r02_bench_cartpole_ps.json 2> gpurun_out/r02_bench_cartpole_ps.err
( timeout 300 python bench.py --workload fr3_pick_cem --steps 5 --warmup 3 --no-extras ) > gpurun_out/r02_bench_fr3_pick_cem.json 2> gpurun_out/r02_bench_fr3_pick_cem.err
( B200MPC_LEAP_PROF=1 timeout 300 python bench.py --workload leap_cube_mppi --steps 3 --warmup 3 --no-extras ) > /dev/null 2> gpurun_out/r02_leap_prof_final.txt
python - <<'PY'
import json, glob
d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default: value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'p50', d['e2e'].get('plan_latency_p50_ms'), 'c1 p50', d.get('plan_latency_c1_p50_ms'))
for k, v in (d.get('also') or {}).items():
    print('also', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'contact_overflows', 'error')}, (v.get('e2e') or {}).get('plan_latency_p50_ms'), v['roofline'].get('frac'))
for f in sorted(glob.glob('gpurun_out/r02_bench_reference_*.json')) + ['gpurun_out/r02_bench_cartpole_ps.json', 'gpurun_out/r02_bench_fr3_pick_cem.json']:
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', r['value'], 'ms/step', r['ms_per_step'], 'p50', r.get('plan_latency_p50_ms'), (r.get('cpu_baseline') or {}).get('cores'))
    except Exception as e:
        print(f, 'failed', e)
PY
