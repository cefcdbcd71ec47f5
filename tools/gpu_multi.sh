#!/bin/bash
# 2-GPU session: multi-GPU tests (peer exchange + all_gather paths), fr3 / leap / cartpole 2-GPU bench lines
mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -8 ) > gpurun_out/pytest_multi.log 2>&1
for wl in cartpole_mppi leap_cube_mppi fr3_pick_cem; do
  ( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --workload $wl --steps 20 --warmup 5 ) > gpurun_out/bench2_$wl.json 2> gpurun_out/bench2_$wl.err
done
cat gpurun_out/pytest_multi.log
python - <<'PY'
import json
for wl in ('cartpole_mppi', 'leap_cube_mppi', 'fr3_pick_cem'):
    try:
        d = json.loads(open(f'gpurun_out/bench2_{wl}.json').read().strip().splitlines()[-1])
        print(wl, 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', round(d['value']), d['config'].get('exchange_used'))
    except Exception as e:
        print(wl, 'failed', e, open(f'gpurun_out/bench2_{wl}.err').read()[-400:])
PY
