#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2y_2gpu.json 2> gpurun_out/r2y_2gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r2y_2gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2y_2gpu.json').read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','n_gpus','steps','gpu_launches','frozen_campaigns','parity')}, 'e2e', round(d['e2e']['value']), 'cpu', d['cpu_baseline']['value'])
PY
