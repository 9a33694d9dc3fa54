#!/bin/bash
# Rebuilds the library with different min-blocks-per-SM for k_var_fit and times it (run on the GPU box).
for mb in 1 2 3 4; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DSTP_VAR_MINB=$mb \
       -Xptxas -v -o stpbenchmark_b200/libstp_b200.so stpbenchmark_b200/csrc/stp_kernels.cu 2>&1 | grep -A2 "k_var_fit\|10VarFitArgsEiNS" | grep -E "registers|spill" | head -2
  echo "== STP_VAR_MINB=$mb"
  python tools/gpu_time_var.py 100 2>&1 | grep -E "B=  296 n= 50|B=  296 n= 89|B= 1024"
done
