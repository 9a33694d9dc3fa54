#!/bin/bash
mkdir -p gpurun_out
for v in base t2c8 rc32; do
  echo "== $v"
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c20_${v}_phases_cfg5.json 2> gpurun_out/r2c20_${v}.err; echo rc=$?
  STP_LIB=$PWD/tools/ab/libstp_$v.so python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -2
done
python bench.py > gpurun_out/r2c20_exact.json 2> gpurun_out/r2c20_exact.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c20_exact.json').read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','steps','gpu_launches','frozen_campaigns','setup_s','parity')}, 'frac', d['roofline']['frac'], 'e2e', d['e2e'])
PY
python tools/gpu_profile_e2e.py > gpurun_out/r2c20_e2e_profile.txt 2>&1; head -24 gpurun_out/r2c20_e2e_profile.txt | cut -c1-150
