#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-extras --steps 200 --warmup 20 > gpurun_out/r02_bench_c2_straight.json 2> gpurun_out/r02_bench_c2_straight.err
timeout 300 python bench.py --workload cylinder_push_cem --no-extras --steps 200 --warmup 20 > gpurun_out/r02_bench_c3_straight.json 2> gpurun_out/r02_bench_c3_straight.err
python -c "
import json
for f in ('c2','c3'):
    d=json.loads(open('gpurun_out/r02_bench_%s_straight.json'%f).read().strip().splitlines()[-1])
    print(f, 'ms/step', d['ms_per_step'], 'value', d['value'], 'kernel_ms', d['roofline']['kernel_ms'])
"
