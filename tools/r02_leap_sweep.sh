#!/bin/bash
# round 2: barrier-period sweep of the leap kernel (sync mode 1 = block barrier per `period` time steps, no lock-step Newton)
mkdir -p gpurun_out
run() { tag=$1; shift; ( env "$@" timeout 120 python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 0.5 --no-extras ) > gpurun_out/r02_leap_$tag.json 2> gpurun_out/r02_leap_$tag.err; }
run base X=1
for p in 1 2 4 8 40; do run sync1_p$p B200MPC_LEAP_SYNC=1 B200MPC_LEAP_SYNC_PERIOD=$p; done
run sync0 B200MPC_LEAP_SYNC=0
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02_leap_*.json')):
    try:
        d = json.load(open(f)); print(f, 'ms/step', round(d['ms_per_step'], 3), 'kernel', round(d['roofline']['kernel_ms'], 3), 'overflows', d.get('contact_overflows'))
    except Exception as e:
        print(f, 'failed', e)
PY
