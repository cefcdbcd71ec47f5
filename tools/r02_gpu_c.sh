#!/bin/bash
mkdir -p gpurun_out
( B200MPC_LEAP_PROF=1 timeout 300 python bench.py --workload leap_cube_mppi --steps 3 --warmup 3 --no-extras ) > gpurun_out/r02_leap_prof_c0.json 2> gpurun_out/r02_leap_prof_c0.txt
grep "leap_blk\|leap_prof" gpurun_out/r02_leap_prof_c0.txt | head -34
( timeout 300 python bench.py --workload leap_cube_mppi --steps 10 --warmup 3 --no-extras ) > gpurun_out/r02_bench_leap_hh.json 2> gpurun_out/r02_bench_leap_hh.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_leap_hh.json').read().strip().splitlines()[-1])
print('leap ms/step', d['ms_per_step'], 'kernel_ms', d['roofline']['kernel_ms'], 'overflows', d.get('contact_overflows'))
PY
