#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_exact --launch-skip 160 -c 200 --csv --log-file gpurun_out/r2f_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_ncu_launch.log 2>&1; tail -2 gpurun_out/r2f_launches.csv | cut -c1-200
python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_exact_steps12.json 2>/dev/null; cut -c1-160 gpurun_out/r2f_exact_steps12.json
python bench.py --groups 1 --schedule mixed --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2f_exact_groups1.json 2>/dev/null; cut -c1-160 gpurun_out/r2f_exact_groups1.json
