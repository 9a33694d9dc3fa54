#!/bin/bash
# ncu launch list of the default device-resident arm (24 cohorts), after the same command has run without ncu
timeout 60 python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_exact_steps12.json 2> /dev/null; cut -c1-150 gpurun_out/r2w_exact_steps12.json
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_exact --launch-skip 96 -c 288 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 12 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r2w_ncu_launch.log 2>&1; tail -1 gpurun_out/r2w_launches.csv | cut -c1-200
