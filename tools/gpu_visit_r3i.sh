#!/bin/bash
# ncu --set full of back_kernel and fwd_kernel in the final build (C2 batch)
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'back_kernel|fwd_kernel' -s 8 -c 2 \
   -o gpurun_out/prof_r02_seed_final -f python bench.py --steps 1 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline --no-chain > gpurun_out/prof_seed_final.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r02_seed_final.ncu-rep
