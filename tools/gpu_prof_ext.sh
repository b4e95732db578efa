#!/bin/bash
# ncu --set full of the hot kernels of one bench step (after 3 warm-up steps) for both extension variants,
# the INT microbenchmark and an A/B bench without the two-row kernel.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
./tools/ubench_int.bin > gpurun_out/ubench_int.txt 2>&1; cat gpurun_out/ubench_int.txt
# matched kernels per step: fwd, back, locate, cut + 6 ext_simd bins = 10; skip the 3 warm-up steps
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'back_kernel|fwd_kernel|ext_simd_kernel|locate_kernel|cut_kernel' -s 30 -c 10 \
   -o gpurun_out/prof_r01_simd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_simd.log 2>&1; echo "ncu simd rc=$?"
# 32-bit kernel only: 7 ext_inter bins per step; capture q64 and q128 (3rd and 4th)
BWA_B200_EXT_NO_SIMD=1 timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'ext_inter_kernel' -s 23 -c 2 \
   -o gpurun_out/prof_r01_inter -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_inter.log 2>&1; echo "ncu inter rc=$?"
BWA_B200_EXT_NO_SIMD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nosimd.json 2> gpurun_out/bench_nosimd.err; echo "bench nosimd rc=$?"
ls -la gpurun_out/
