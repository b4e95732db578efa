#!/bin/bash
# ncu --set full of the pair-kernel bins of one bench step (after 3 warm-up steps)
set -u
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'ext_pair_kernel' -s ${SKIP:-40} -c ${COUNT:-3} \
   -o gpurun_out/prof_r01_pair -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pair.log 2>&1; echo "ncu pair rc=$?"
