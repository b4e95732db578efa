#!/bin/bash
set -u
mkdir -p gpurun_out
BWA_B200_FWD_MINB=${FM:-8} BWA_B200_BACK_MINB=${BM:-8} timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:"${KREGEX:-back_kernel|fwd_kernel}" -s ${SKIP:-6} -c ${COUNT:-2} \
   -o gpurun_out/${OUT:-prof_seed} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof.log 2>&1; echo "ncu rc=$?"
tail -2 gpurun_out/prof.log | cut -c1-300
