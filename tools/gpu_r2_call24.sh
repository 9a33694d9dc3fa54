#!/bin/bash
mkdir -p gpurun_out
for v in head tb2; do
  echo "== $v"
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c24_${v}_phases_cfg5.json 2> gpurun_out/r2c24_${v}.err; echo rc=$?
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -2
done
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -2
bash tools/gpu_r2_call23.sh
ncu --metrics l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__sass_inst_executed_op_shared_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum ./tools/micro/lds64_probe.bin > gpurun_out/r2c24_lds64_probe.txt 2>&1; grep -v "^==PROF==" gpurun_out/r2c24_lds64_probe.txt | tail -20
