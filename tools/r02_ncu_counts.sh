#!/bin/bash
# Per-launch hardware counts of the dominant kernels (DRAM bytes, warp / thread / fp64 instructions, cycles) -> gpurun_out/r02_counts_*.csv,
# turned into profiles/r02_counts.json by tools/r02_make_counts.py.  bench.py divides them by the kernel time it measures live.
# Launch order of `bench.py --no-extras --steps 3 --warmup 3`: 3 warm-up + 3 timed plan steps, then 3 launches of the rollout+cost kernel alone
# (the roofline kernel): -s 6 -c 1 captures the first of those.
mkdir -p gpurun_out
M=dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_fp64.sum,sm__pipe_fp64_cycles_active.avg,sm__cycles_elapsed.max,gpu__time_duration.sum,sm__warps_active.avg.per_cycle_active,smsp__issue_active.avg.pct_of_peak_sustained_active
for w in cartpole_mppi:rollout_kernel cylinder_push_cem:rollout_kernel leap_cube_mppi:leap_rollout_kernel ${EXTRA_WORKLOADS}; do
  name=${w%%:*}; kern=${w##*:}
  timeout 300 ncu --metrics $M --clock-control none -k regex:$kern -s 6 -c 1 --csv --log-file gpurun_out/r02_counts_$name.csv \
    python bench.py --workload $name --steps 3 --warmup 3 --no-extras > /dev/null 2> gpurun_out/r02_counts_$name.err
done
if [ -n "$FULL" ]; then
  for w in $FULL; do
    name=${w%%:*}; kern=${w##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s 6 -c 1 -f -o gpurun_out/r02_full_$name \
      python bench.py --workload $name --steps 3 --warmup 3 --no-extras > /dev/null 2> gpurun_out/r02_full_$name.err
  done
fi
ls -la gpurun_out/r02_counts_*.csv
