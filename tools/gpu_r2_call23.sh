#!/bin/bash
# e2e tuning of the product path: sub-group size x chunk
mkdir -p gpurun_out
for mg in 64 128 256 512; do for ch in 8 16; do
  echo -n "max_group=$mg chunk=$ch: "; STP_JOB_MAX_GROUP=$mg STP_JOB_CHUNK=$ch python tools/gpu_profile_e2e.py 2>&1 | grep -m2 "run_sweep(\|Event.synchronize" | tr '\n' ' ' | cut -c1-230; echo
done; done
