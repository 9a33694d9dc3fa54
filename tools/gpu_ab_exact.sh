#!/bin/bash
# A/B of exact-kernel thread counts on the GPU box (rebuilds the library in place).
run() {  # $1 = nvcc defines, $2 = STP_THREADS (or empty)
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $1 -Xptxas -v \
       -o stpbenchmark_b200/libstp_b200.so stpbenchmark_b200/csrc/stp_kernels.cu 2>&1 | grep -A1 "bo_step" | grep -E "spill" | head -1
  echo "== $1 threads=$2"
  STP_THREADS=$2 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['frozen_campaigns'])"
}
run "-DSTP_EXACT_MAXT=320 -DSTP_EXACT_MINB=2" "320"
run "-DSTP_EXACT_MAXT=384 -DSTP_EXACT_MINB=2" "384"
run "-DSTP_EXACT_MAXT=512 -DSTP_EXACT_MINB=2" "512"
run "-DSTP_EXACT_MAXT=512 -DSTP_EXACT_MINB=1" "512"
