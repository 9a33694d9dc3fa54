#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -3
for v in nostage t2c4 t3c6; do
  echo "== $v"
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_phase_timers_var.py > gpurun_out/r2c18_${v}_phases.json 2> gpurun_out/r2c18_${v}.err; echo rc=$?
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c18_${v}_phases_cfg5.json 2>> gpurun_out/r2c18_${v}.err; echo rc=$?
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_time_var.py 40 2>&1 | tail -6
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -2
done
python tools/gpu_profile_e2e.py > gpurun_out/r2c18_e2e_profile.txt 2>&1; head -60 gpurun_out/r2c18_e2e_profile.txt
