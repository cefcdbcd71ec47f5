#!/bin/bash
mkdir -p gpurun_out
for hh in 1 0; do
  ( B200MPC_LEAP_NO_HH=$hh timeout 300 python bench.py --workload leap_cube_mppi --steps 10 --warmup 3 --no-extras ) > gpurun_out/r02_bench_leap_nohh$hh.json 2> gpurun_out/r02_bench_leap_nohh$hh.err
  python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_leap_nohh$hh.json').read().strip().splitlines()[-1])
print('NO_HH=$hh leap ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'overflows', d.get('contact_overflows'))
PY
done
( B200MPC_LEAP_NO_HH=1 B200MPC_LEAP_PROF=1 timeout 300 python bench.py --workload leap_cube_mppi --steps 3 --warmup 3 --no-extras ) > /dev/null 2> gpurun_out/r02_leap_prof_nohh.txt
grep leap_prof gpurun_out/r02_leap_prof_nohh.txt | head -19
