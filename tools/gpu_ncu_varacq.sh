#!/bin/bash
# ncu --set full of the Monte-Carlo pool pass at both CTA shapes (n = 50: 256 threads x 3 CTAs/SM; n = 89: 512 threads x 1)
mkdir -p gpurun_out
for n in 50 89; do
ncu --set full --clock-control none -k regex:k_var_acq -s 1 -c 1 -f -o gpurun_out/prof_r2f_varacq_n$n python tools/gpu_prof_acq.py $n > gpurun_out/r2f_ncu_varacq_n$n.log 2>&1; tail -1 gpurun_out/r2f_ncu_varacq_n$n.log
done
