#!/bin/bash
# One GPU-box session for the fr3_pick bring-up: parity tests, bench lines, ncu launch list + full capture of the fr3 kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/gpu.txt 2>&1
( timeout 300 python -m pytest tests/test_gpu_fr3.py -q 2>&1 | tail -60 ) > gpurun_out/pytest_fr3.log 2>&1
( timeout 200 python bench.py --workload fr3_pick_cem --steps 10 --warmup 3 --cpu-budget 6 ) > gpurun_out/bench_fr3.json 2> gpurun_out/bench_fr3.err
( timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fr3.csv python bench.py --workload fr3_pick_cem --steps 2 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_launch.log 2>&1
( timeout 240 ncu --set full --clock-control none --import-source on -k regex:fr3_rollout_kernel -c 1 -o gpurun_out/fr3_full python bench.py --workload fr3_pick_cem --steps 1 --warmup 1 --cpu-budget 1 ) > gpurun_out/ncu_full.log 2>&1
( timeout 100 ncu -i gpurun_out/fr3_full.ncu-rep --page raw --csv ) > gpurun_out/fr3_full_raw.csv 2>&1
( timeout 100 python bench.py --workload fr3_pick_cem --n-rollouts 64 --steps 10 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_n64.json 2> gpurun_out/bench_fr3_n64.err
( timeout 100 python bench.py --workload fr3_pick_cem --n-rollouts 4096 --steps 5 --warmup 3 --cpu-budget 1 ) > gpurun_out/bench_fr3_n4096.json 2> gpurun_out/bench_fr3_n4096.err
( timeout 400 python -m pytest tests -m gpu -q --deselect tests/test_gpu_fr3.py 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log 2>&1
( timeout 120 python bench.py --steps 200 --warmup 20 --cpu-budget 4 ) > gpurun_out/bench_cartpole.json 2> gpurun_out/bench_cartpole.err
( timeout 120 python bench.py --workload cylinder_push_cem --steps 200 --warmup 20 --cpu-budget 4 ) > gpurun_out/bench_cyl.json 2> gpurun_out/bench_cyl.err
( timeout 120 python bench.py --workload leap_cube_mppi --steps 20 --warmup 3 --cpu-budget 4 ) > gpurun_out/bench_leap.json 2> gpurun_out/bench_leap.err
ls -la gpurun_out
tail -5 gpurun_out/pytest_fr3.log gpurun_out/pytest_gpu.log
cat gpurun_out/bench_fr3.json
