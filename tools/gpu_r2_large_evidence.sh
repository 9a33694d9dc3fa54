#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_large.py -q -x 2>&1 | tail -4
timeout 600 python bench.py --config f3 > gpurun_out/r2_bench_f3.json 2> gpurun_out/r2_bench_f3.err
cut -c1-600 gpurun_out/r2_bench_f3.json; tail -3 gpurun_out/r2_bench_f3.err
timeout 300 python tools/gpu_single_large.py ExactGP && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_exact_fit -c 1 -f -o gpurun_out/prof_r2_exact_fit_large \
  python tools/gpu_single_large.py ExactGP > gpurun_out/ncu_r2_exact_fit_large.log 2>&1
tail -3 gpurun_out/ncu_r2_exact_fit_large.log
