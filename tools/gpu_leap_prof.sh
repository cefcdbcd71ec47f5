#!/bin/bash
mkdir -p gpurun_out
( timeout 300 ncu --set full --clock-control none --import-source on -k regex:leap_rollout_kernel -c 1 -f -o gpurun_out/leap_full3 python bench.py --workload leap_cube_mppi --steps 1 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_leap.log 2>&1
( B200MPC_LEAP_PROF=1 timeout 100 python bench.py --workload leap_cube_mppi --steps 3 --warmup 1 --cpu-budget 1 ) > gpurun_out/prof_leap.json 2> gpurun_out/prof_leap.err
grep leap_prof gpurun_out/prof_leap.err
ls -la gpurun_out/leap_full3.ncu-rep
