#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_exact.py -m gpu -x -q -k "one_device_job or cfg1 or persistent" 2>&1 | tail -2
python bench.py --no-cpu-baseline > gpurun_out/r2c49_exact.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r2c49_exact.json').read().strip().splitlines()[-1])
print('exact', round(d['value'],1), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'})
PY
python tools/gpu_profile_e2e.py 2>&1 | head -16 | cut -c1-150
