#!/bin/bash
# ncu --set full capture of the hot kernels of one bench step (after 3 warm-up steps), plus the INT microbenchmark.
set -u
mkdir -p gpurun_out
./tools/ubench_int.bin > gpurun_out/ubench_int.txt 2>&1; cat gpurun_out/ubench_int.txt
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'back_kernel|fwd_kernel|ext_inter_kernel|locate_kernel|cut_kernel' -s ${SKIP:-33} -c ${COUNT:-11} \
   -o gpurun_out/prof_r01 -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/prof.log
ls -la gpurun_out/
