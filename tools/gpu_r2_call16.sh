#!/bin/bash
mkdir -p gpurun_out
./tools/micro/dmma_probe.bin > gpurun_out/r2c16_dmma_probe.txt 2>&1; cat gpurun_out/r2c16_dmma_probe.txt
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -4
python tools/gpu_phase_timers_var.py > gpurun_out/r2c16_var_phases.json 2> gpurun_out/r2c16_var_phases.err; echo rc=$?
python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c16_var_phases_cfg5.json 2>> gpurun_out/r2c16_var_phases.err; echo rc=$?
python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -3
