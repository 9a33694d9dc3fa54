#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do for v in head gram both; do
  STP_LIB=$PWD/tools/ab/libstp_$v.so python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2c29_${v}_$rep.json 2>/dev/null
  python - <<PY
import json
d=json.loads(open('gpurun_out/r2c29_${v}_$rep.json').read().strip().splitlines()[-1])
print('$v rep $rep', round(d['value'],1), round(d['ms_per_step'],3), d['parity'] if 'parity' in d else '')
PY
done; done
STP_LIB=$PWD/tools/ab/libstp_both.so python -m pytest tests/test_gpu_exact.py tests/test_gpu_determinism.py -m gpu -x -q 2>&1 | tail -2
