#!/bin/bash
for z in 0 1; do
  for wl in cartpole_mppi cylinder_push_cem leap_cube_mppi; do
    B200MPC_ZEROCOPY=$z python bench.py --workload $wl --steps 100 --warmup 10 --cpu-budget 0.5 2>/dev/null > /tmp/o.json
    python -c "import json; d=json.loads(open('/tmp/o.json').read().strip().splitlines()[-1]); print('zerocopy $z $wl', 'resident ms', round(d['ms_per_step'],4), 'e2e p50 ms', round(d['e2e']['plan_latency_p50_ms'],4), 'e2e rollouts/s', round(d['e2e']['value']))"
  done
done
