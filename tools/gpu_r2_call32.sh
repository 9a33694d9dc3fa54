#!/bin/bash
mkdir -p gpurun_out
for v in u3 u6 u8; do
  echo "== $v"
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_phase_timers_var.py > gpurun_out/r2c32_${v}_phases.json 2> gpurun_out/r2c32_${v}.err; echo rc=$?
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_time_var.py 40 2>&1 | tail -6
done
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -2
