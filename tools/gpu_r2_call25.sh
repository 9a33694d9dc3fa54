#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_var.py -m gpu -x -q 2>&1 | tail -3
python tools/gpu_phase_timers_var.py cfg5 > gpurun_out/r2c25_phases_cfg5.json 2> gpurun_out/r2c25.err; echo rc=$?
python tools/gpu_time_var.py 10 cfg5 2>&1 | tail -2
python bench.py --config cfg5 --no-cpu-baseline > gpurun_out/r2c25_cfg5.json 2> gpurun_out/r2c25_cfg5.err; echo "rc=$?"; cut -c1-200 gpurun_out/r2c25_cfg5.json
python bench.py --steps 80 --no-cpu-baseline > gpurun_out/r2c25_exact.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r2c25_exact.json').read().strip().splitlines()[-1])
print('exact s80', d['value'], {k:v for k,v in d['e2e'].items() if k!='what'})
PY
