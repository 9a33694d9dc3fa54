#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_parity_sweep.py --out gpurun_out/r02_parity_gold1.json 2>&1 | tail -30
