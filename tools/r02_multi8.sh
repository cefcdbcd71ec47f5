#!/bin/bash
# 8-GPU session: scaling bench at 8 ranks (C2 in-kernel exchange + also C5 = leap_cube 8 x 1024) and C5 through the plugin surface
N=${1:-8}
mkdir -p gpurun_out
( timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 100 --warmup 10 ) > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err
tail -2 gpurun_out/r02_bench_${N}gpu.err
python - <<PY
import json
d = json.loads(open('gpurun_out/r02_bench_${N}gpu.json').read().strip().splitlines()[-1])
print('n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 4), 'value', round(d['value']), d.get('exchange_used'), d.get('exchange_verified'))
print('exchange_timing', d.get('exchange_timing'))
for k, v in (d.get('also') or {}).items():
    print('also', k, {kk: v.get(kk) for kk in ('value', 'ms_per_step', 'exchange_used', 'exchange_verified', 'contact_overflows', 'error')})
PY
python - <<PY
import sys, time, statistics, numpy as np
sys.path.insert(0, '.')
from judo_b200.controller import make_controller
for task, opt, n, hor in (('leap_cube', 'mppi', 1024 * $N, 0.4),):
    np.random.seed(42)
    c = make_controller(task, opt, devices=list(range($N)))
    c.optimizer_cfg.num_rollouts = n
    c.controller_cfg.horizon = hor
    c.reset()
    if hasattr(c.task, 'get_sim_metadata'): c.system_metadata = c.task.get_sim_metadata()
    for _ in range(3): c.update_action()
    lat = []
    for i in range(10):
        c.time = c.task.dt * i
        t = time.perf_counter(); c.update_action(); lat.append(time.perf_counter() - t)
    print(f'controller devices=$N {task}+{opt} N={n}: p50 {statistics.median(lat)*1e3:.3f} ms -> {n/statistics.median(lat):.0f} rollouts/s; overflows {c.engine.contact_overflows}; fast {c._can_fast_path()}')
    c.engine.close()
PY
