#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -2
for g in 2 4 8; do
python bench.py --schedule queue --groups $g --no-cpu-baseline --no-e2e > gpurun_out/r2c21_queue_g$g.json 2> gpurun_out/r2c21_queue_g$g.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c21_queue_g$g.json').read().strip().splitlines()[-1])
print('queue groups $g', {k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','steps','gpu_launches','frozen_campaigns')}, 'frac', d['roofline']['frac'])
PY
done
python bench.py --schedule queue --groups 4 --steps 80 --no-cpu-baseline --no-e2e > gpurun_out/r2c21_queue_g4_s80.json 2> gpurun_out/r2c21_queue_g4_s80.err; cut -c1-150 gpurun_out/r2c21_queue_g4_s80.json
python tools/gpu_profile_e2e.py > gpurun_out/r2c21_e2e_profile.txt 2>&1; head -16 gpurun_out/r2c21_e2e_profile.txt | cut -c1-150
python bench.py --config cfg5 --no-cpu-baseline > gpurun_out/r2c21_cfg5.json 2> gpurun_out/r2c21_cfg5.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2c21_cfg5.json
