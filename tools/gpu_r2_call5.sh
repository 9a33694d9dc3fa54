#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py -m gpu -x -q -s -k "known_answer or teacher_forced_along" 2>&1 | grep -v "^$" | tail -8
python tools/gpu_single_n.py --n 60 --B 296
ncu --set full --clock-control none --import-source on -k regex:k_exact_bo_step -s 1 -c 1 -f -o gpurun_out/prof_r2_bostep_n60 python tools/gpu_single_n.py --n 60 --B 296 > gpurun_out/ncu_r2_n60.log 2>&1; tail -2 gpurun_out/ncu_r2_n60.log
python -m pytest tests -m gpu -q 2>&1 | tail -5
