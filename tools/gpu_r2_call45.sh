#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py tests/test_gpu_hartmann.py -m gpu -x -q 2>&1 | tail -12
python tools/gpu_time_acq_split.py 2>&1 | tail -8
python tools/gpu_time_var.py 40 2>&1 | tail -6
python bench.py --model VarTGP --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c45_vartgp.json 2>/dev/null; cut -c1-120 gpurun_out/r2c45_vartgp.json
python bench.py --config cfg2 --no-cpu-baseline > gpurun_out/r2c45_cfg2.json 2>/dev/null; cut -c1-120 gpurun_out/r2c45_cfg2.json
