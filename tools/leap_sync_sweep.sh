#!/bin/bash
# compare the block-barrier modes of the leap kernel (0 none, 1 per step, 2 per step + mid-step)
for m in 2 3; do
  B200MPC_LEAP_SYNC=$m python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 1 2>/dev/null > /tmp/o.json
  python -c "import json; d=json.loads(open('/tmp/o.json').read().strip().splitlines()[-1]); print('sync mode $m', d['ms_per_step'], d['value'])"
done
