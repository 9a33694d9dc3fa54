#!/bin/bash
# usage: bash tools/gpu_ngpu_check.sh N  -- smoke() + the default bench arm and its reference arm under torchrun at N ranks
N=${1:-2}
mkdir -p gpurun_out
P=29531
python __graft_entry__.py --smoke 2>&1 | tail -2
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --impl reference --steps 4 --warmup 3 > gpurun_out/r2c30_ref_${N}gpu.json 2> gpurun_out/r2c30_ref_${N}gpu.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/r2c30_ref_${N}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 40 --warmup 3 > gpurun_out/r2c30_bench_${N}gpu.json 2> gpurun_out/r2c30_bench_${N}gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r2c30_bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c30_bench_${N}gpu.json').read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','n_gpus','steps','gpu_launches','frozen_campaigns','setup_s','parity')}, 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'}, 'cpu', d['cpu_baseline']['value'])
PY
