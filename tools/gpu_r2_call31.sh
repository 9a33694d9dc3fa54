#!/bin/bash
mkdir -p gpurun_out
python tools/gpu_time_acq_split.py 2>&1 | tee gpurun_out/r2c31_acq_split.txt | tail -10
