#!/bin/bash
# Saturation curve of the thread-per-rollout kernel ("nobody has measured the saturation curve that would state the card's
# real rollouts/s"): cartpole+mppi H=64, N = 4 K ... 1 M per launch, resident plan step (CUDA events, L2 flushed).
mkdir -p gpurun_out
for n in 4096 16384 65536 262144 1048576; do
  ( timeout 200 python bench.py --workload cartpole_mppi --n-rollouts $n --steps 30 --warmup 5 --no-extras ) > gpurun_out/r02_sat_cartpole_$n.json 2> gpurun_out/r02_sat_cartpole_$n.err
done
for n in 2048 65536 524288; do
  ( timeout 200 python bench.py --workload cylinder_push_cem --n-rollouts $n --steps 30 --warmup 5 --no-extras ) > gpurun_out/r02_sat_cylinder_$n.json 2> gpurun_out/r02_sat_cylinder_$n.err
done
for n in 1024 2048 4096 8192; do
  ( timeout 300 python bench.py --workload leap_cube_mppi --n-rollouts $n --steps 5 --warmup 3 --no-extras ) > gpurun_out/r02_sat_leap_$n.json 2> gpurun_out/r02_sat_leap_$n.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob('gpurun_out/r02_sat_*.json'), key=lambda p: (p.split('_')[-2], int(p.split('_')[-1][:-5]))):
    try:
        d = json.load(open(f)); print(f, 'ms/step', round(d['ms_per_step'], 4), 'rollouts/s', f"{d['value']:.4g}", 'state-steps/s', f"{d['state_steps_per_s']:.4g}", 'kernel_ms', round(d['roofline']['kernel_ms'], 4), 'hbm_frac', f"{d['roofline'].get('frac', 0):.3g}")
    except Exception as e:
        print(f, 'failed', e)
PY
