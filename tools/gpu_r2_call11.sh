#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python bench.py > gpurun_out/r2c11_exact.json 2> gpurun_out/r2c11_exact.err; echo "rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c11_exact.json').read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','steps','gpu_launches','frozen_campaigns','setup_s','parity')}, 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'])
PY
