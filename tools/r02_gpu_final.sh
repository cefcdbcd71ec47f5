#!/bin/bash
# final one-GPU pass of round 2: ncu counts (-> profiles/r02_counts.json, read by bench.py), GPU test tier, smoke, default bench line
# (C2 + also C3/C4), C1 / fr3 lines, leap phase timers, launch list of the default step
mkdir -p gpurun_out
EXTRA_WORKLOADS=fr3_pick_cem:fr3_rollout_kernel bash tools/r02_ncu_counts.sh > /dev/null 2>&1
python tools/r02_make_counts.py > /dev/null
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6 ) > gpurun_out/r02_pytest_gpu.log 2>&1
cat gpurun_out/r02_pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 ) > gpurun_out/r02_smoke.log; cat gpurun_out/r02_smoke.log
( timeout 600 python bench.py ) > gpurun_out/r02_bench_default.json 2> gpurun_out/r02_bench_default.err
tail -2 gpurun_out/r02_bench_default.err
( timeout 300 python bench.py --workload cartpole_ps --steps 100 --warmup 10 --no-extras ) > gpurun_out/r02_bench_cartpole_ps.json 2> gpurun_out/r02_bench_cartpole_ps.err
( timeout 300 python bench.py --workload fr3_pick_cem --steps 5 --warmup 3 --no-extras ) > gpurun_out/r02_bench_fr3_pick_cem.json 2> gpurun_out/r02_bench_fr3_pick_cem.err
( B200MPC_LEAP_PROF=1 timeout 300 python bench.py --workload leap_cube_mppi --steps 3 --warmup 3 --no-extras ) > /dev/null 2> gpurun_out/r02_leap_prof_final.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_cartpole_mppi.csv python bench.py --steps 5 --warmup 3 --no-extras > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches_leap_cube_mppi.csv python bench.py --workload leap_cube_mppi --steps 3 --warmup 3 --no-extras > /dev/null 2>&1
python tools/r02_latency.py > gpurun_out/r02_latency.txt 2>&1
python - <<'PY'
import json, glob
d = json.loads(open('gpurun_out/r02_bench_default.json').read().strip().splitlines()[-1])
print('default: value', d['value'], 'ms/step', d['ms_per_step'], 'e2e', d['e2e']['value'], 'p50', d['e2e'].get('plan_latency_p50_ms'), 'c1 p50', d.get('plan_latency_c1_p50_ms'), 'roofline', d['roofline'].get('frac'), d['roofline'].get('fp64_issue', {}).get('frac'))
for k, v in (d.get('also') or {}).items():
    print('also', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'contact_overflows', 'error')}, (v.get('e2e') or {}).get('plan_latency_p50_ms'), v['roofline'].get('frac'))
for f in ['gpurun_out/r02_bench_cartpole_ps.json', 'gpurun_out/r02_bench_fr3_pick_cem.json']:
    try:
        r = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, 'value', r['value'], 'ms/step', r['ms_per_step'], 'p50', r.get('plan_latency_p50_ms'))
    except Exception as e:
        print(f, 'failed', e)
PY
