#!/bin/bash
# same-box A/B of library variants (gpurun_variants_*.so): two rounds each; WORKLOAD selects the bench workload (default C2)
W=${WORKLOAD:-cartpole_mppi}; S=${STEPS:-300}
mkdir -p gpurun_out
cp judo_b200/libb200mpc.so /tmp/lib_orig.so
for round in 1 2; do
for f in gpurun_variants_*.so; do
  cp $f judo_b200/libb200mpc.so
  timeout 300 python bench.py --workload $W --no-extras --steps $S --warmup 5 > gpurun_out/ab.json 2> gpurun_out/ab.err
  python -c "
import json
d=json.loads(open('gpurun_out/ab.json').read().strip().splitlines()[-1])
print('$f round $round', 'ms/step %.5f' % d['ms_per_step'], 'kernel_ms %.5f' % d['roofline']['kernel_ms'], d.get('contact_overflows'))
"
done
done
cp /tmp/lib_orig.so judo_b200/libb200mpc.so
