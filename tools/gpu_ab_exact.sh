#!/bin/bash
# A/B of the exact closure variants on the GPU box (rebuilds the library in place).
for cfg in "-DSTP_EXACT_REGS=1 -DSTP_REG_BS=6 -DSTP_EXACT_MAXT=160" "-DSTP_EXACT_REGS=1 -DSTP_REG_BS=5 -DSTP_EXACT_MAXT=192" "-DSTP_EXACT_REGS=1 -DSTP_REG_BS=8 -DSTP_EXACT_MAXT=128 -DSTP_EXACT_MINB=2"; do
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $cfg \
       -o stpbenchmark_b200/libstp_b200.so stpbenchmark_b200/csrc/stp_kernels.cu 2>&1 | grep -i error
  echo "== $cfg"
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['parity'], d['frozen_campaigns'])"
done
