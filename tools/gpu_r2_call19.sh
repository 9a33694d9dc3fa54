#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c19_var_phases_cfg5.json 2> gpurun_out/r2c19_var_phases.err; echo rc=$?
python tools/gpu_time_var.py 40 2>&1 | tail -6
python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -2
python bench.py --model VarTGP --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c19_vartgp.json 2> gpurun_out/r2c19_vartgp.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2c19_vartgp.json
python bench.py --config cfg5 --no-cpu-baseline > gpurun_out/r2c19_cfg5.json 2> gpurun_out/r2c19_cfg5.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2c19_cfg5.json
