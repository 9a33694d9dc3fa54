#!/bin/bash
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_var_acq -s 1 -c 1 -f -o gpurun_out/prof_r2f_varacq_n50 python tools/gpu_prof_acq.py > gpurun_out/r2f_ncu_varacq.log 2>&1; tail -1 gpurun_out/r2f_ncu_varacq.log
python bench.py --no-cpu-baseline > gpurun_out/r2c44_exact.json 2>/dev/null; python - <<PY
import json
d=json.loads(open('gpurun_out/r2c44_exact.json').read().strip().splitlines()[-1])
print('exact', round(d['value'],1), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'})
PY
