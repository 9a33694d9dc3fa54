#!/bin/bash
mkdir -p gpurun_out
P=29517
for n in 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --impl reference --steps 4 --warmup 1 > gpurun_out/r2c10_ref_${n}gpu.json 2> gpurun_out/r2c10_ref_${n}gpu.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2c10_ref_${n}gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --steps 40 --warmup 3 > gpurun_out/r2c10_bench_${n}gpu.json 2> gpurun_out/r2c10_bench_${n}gpu.err; echo "rc=$?"; tail -c 500 gpurun_out/r2c10_bench_${n}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2c10_bench_${n}gpu.json').read().strip().splitlines()[-1])
print({k:(round(v,2) if isinstance(v,float) else v) for k,v in d.items() if k in ('value','ms_per_step','n_gpus','steps','gpu_launches','frozen_campaigns','setup_s','parity')}, 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='what'}, 'cpu', d['cpu_baseline']['value'])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $P bench.py --gpus $n --config cfg5 --campaigns 148 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r2c10_cfg5_${n}gpu.json 2> gpurun_out/r2c10_cfg5_${n}gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r2c10_cfg5_${n}gpu.err; cut -c1-200 gpurun_out/r2c10_cfg5_${n}gpu.json
done
# the product path under torchrun: the cfg2 sweep (3 seeds, SMOKE length) sharded over the ranks
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --config cfg2 --trials 10 --epochs 100 --no-cpu-baseline > gpurun_out/r2c10_cfg2_2gpu.json 2> gpurun_out/r2c10_cfg2_2gpu.err; echo "rc=$?"; tail -c 300 gpurun_out/r2c10_cfg2_2gpu.err; cut -c1-250 gpurun_out/r2c10_cfg2_2gpu.json
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
