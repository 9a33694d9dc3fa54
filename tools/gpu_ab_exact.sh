#!/bin/bash
# A/B of the exact closure variants on the GPU box (rebuilds the library in place).
for cfg in "0 2" "1 1" "1 2"; do
  set -- $cfg
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC -DSTP_EXACT_REGS=$1 -DSTP_EXACT_MINB=$2 \
       -o stpbenchmark_b200/libstp_b200.so stpbenchmark_b200/csrc/stp_kernels.cu 2>&1 | grep -i error
  echo "== STP_EXACT_REGS=$1 MINB=$2"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'])"
done
