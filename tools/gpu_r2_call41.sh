#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for p in 0 1; do
STP_JOB_PRIORITIES=$p python bench.py --no-cpu-baseline > gpurun_out/r2c41_exact_p${p}_$rep.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r2c41_exact_p${p}_$rep.json').read().strip().splitlines()[-1])
print('prio $p rep $rep: value', round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'wall', round(d['e2e']['wall_s'],3))
PY
done; done
