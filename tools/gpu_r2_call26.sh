#!/bin/bash
mkdir -p gpurun_out
for v in base blk128 blk64x128; do
  echo "== $v"
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c26_${v}_phases_cfg5.json 2> gpurun_out/r2c26_${v}.err; echo rc=$?; tail -2 gpurun_out/r2c26_${v}.err
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -2
done
STP_LIB=$PWD/tools/ab/libstp_blk128.so python -m pytest tests/test_gpu_var.py -m gpu -x -q -k "cfg5" 2>&1 | tail -2
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -2
