#!/bin/bash
# one-GPU session: GPU test tier, leap bench (kernel-only line + phase timers), default bench line
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/r02_pytest_gpu.log 2>&1
cat gpurun_out/r02_pytest_gpu.log
( timeout 300 python bench.py --workload leap_cube_mppi --steps 10 --warmup 3 --no-extras ) > gpurun_out/r02_bench_leap_hh.json 2> gpurun_out/r02_bench_leap_hh.err
( B200MPC_LEAP_PROF=1 timeout 300 python bench.py --workload leap_cube_mppi --steps 3 --warmup 3 --no-extras ) > /dev/null 2> gpurun_out/r02_leap_prof_hh.txt
grep leap_prof gpurun_out/r02_leap_prof_hh.txt
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_leap_hh.json').read().strip().splitlines()[-1])
print('leap ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'overflows', d.get('contact_overflows'))
PY
