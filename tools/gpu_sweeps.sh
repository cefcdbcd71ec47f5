#!/bin/bash
# Experiment knobs of the warp-per-rollout kernels, one bench line per setting (results quoted in profiles/r01c_*.md):
#   B200MPC_LEAP_WPB   warps per block (lock-step group size);  B200MPC_LEAP_SYNC / B200MPC_FR3_SYNC  0 free-running, 1 barrier per time
#   step, 2 + phase barriers, 3 + lock-step Newton iterations (default);  B200MPC_LEAP_PROF / B200MPC_FR3_PROF  clock64 phase timers.
mkdir -p gpurun_out
for w in 7 4 2 1; do
  ( B200MPC_LEAP_WPB=$w timeout 100 python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_leap_wpb$w.json 2> gpurun_out/bench_leap_wpb$w.err
done
for sm in 3 2 1 0; do
  ( B200MPC_LEAP_SYNC=$sm timeout 100 python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_leap_sync$sm.json 2> gpurun_out/bench_leap_sync$sm.err
  ( B200MPC_FR3_SYNC=$sm timeout 100 python bench.py --workload fr3_pick_cem --steps 8 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_sync$sm.json 2> gpurun_out/bench_fr3_sync$sm.err
done
( B200MPC_LEAP_PROF=1 timeout 100 python bench.py --workload leap_cube_mppi --steps 3 --warmup 1 --cpu-budget 1 ) > /dev/null 2> gpurun_out/prof_leap.err
( B200MPC_FR3_PROF=1 timeout 100 python bench.py --workload fr3_pick_cem --steps 3 --warmup 1 --cpu-budget 1 ) > /dev/null 2> gpurun_out/prof_fr3.err
grep -h "_prof " gpurun_out/prof_leap.err gpurun_out/prof_fr3.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/bench_*_wpb*.json') + glob.glob('gpurun_out/bench_*_sync*.json')):
    try:
        d = json.load(open(f))
        print(f, 'ms/step', round(d['ms_per_step'], 3), 'rollouts/s', round(d['value']))
    except Exception as e:
        print(f, 'failed', e)
PY
