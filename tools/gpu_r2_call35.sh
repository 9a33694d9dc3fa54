#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py tests/test_gpu_determinism.py -m gpu -x -q 2>&1 | tail -2
python tools/gpu_time_acq_split.py 2>&1 | tail -8
python bench.py --model VarTGP --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/r2c35_vartgp.json 2>/dev/null; cut -c1-120 gpurun_out/r2c35_vartgp.json
