#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tools/gpu_phase_timers_var.py > gpurun_out/r2c14_var_phases.json 2> gpurun_out/r2c14_var_phases.err; echo rc=$?
python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c14_var_phases_cfg5.json 2>> gpurun_out/r2c14_var_phases.err; echo rc=$?
python tools/gpu_time_var.py 40 2>&1 | tail -8
python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2c14_exact.json 2> gpurun_out/r2c14_exact.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2c14_exact.json
python bench.py --model VarTGP --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c14_vartgp.json 2> gpurun_out/r2c14_vartgp.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2c14_vartgp.json
