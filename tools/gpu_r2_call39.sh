#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py tests/test_gpu_determinism.py tests/test_gpu_hartmann.py -m gpu -x -q 2>&1 | tail -2
python tools/gpu_phase_timers_var.py > gpurun_out/r2c39_phases.json 2> gpurun_out/r2c39.err; echo rc=$?
python tools/gpu_time_var.py 40 2>&1 | tail -6
python bench.py --model VarTGP --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c39_vartgp.json 2>/dev/null; cut -c1-120 gpurun_out/r2c39_vartgp.json
