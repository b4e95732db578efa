#!/bin/bash
# every GPU test on the current library, then ncu --set full of the extension launch set of one chained step (the shipped ext_pair_kernel bins)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ext_pair_kernel' -s 192 -c 16 \
   -o gpurun_out/prof_r02_pair -f python bench.py --steps 1 --warmup 3 --no-extras --no-c3 --no-cpu-baseline > gpurun_out/prof_pair.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/prof_pair.log
