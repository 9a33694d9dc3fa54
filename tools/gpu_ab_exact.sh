#!/bin/bash
# A/B of build-time / launch variants of the exact kernel on the GPU box (rebuilds the library in place; run under
# gpurun).  Each argument is "nvcc defines|STP_THREADS" (either part may be empty); the default build runs last.
run() {
  local defs="${1%%|*}" thr="${1##*|}"
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -shared -Xcompiler -fPIC $defs \
       -o stpbenchmark_b200/libstp_b200.so stpbenchmark_b200/csrc/stp_kernels.cu 2>&1 | grep -E "error" | head -3
  echo "== [$defs] threads=[$thr]"
  for rep in 1 2; do
    STP_THREADS=$thr python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), d['frozen_campaigns'], d['clocks']['sm_mhz'])"
  done
}
for v in "$@"; do run "$v"; done
run "|"
