#!/bin/bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/try.txt
python tools/rs_sweep.py 2>&1 | tee gpurun_out/rs_sweep_fetch64.txt
BWA_B200_L2_FETCH=32 python tools/rs_sweep.py 2>&1 | tee gpurun_out/rs_sweep_fetch32.txt
bash tools/gpu_try.sh "BWA_B200_KMER_K=0 BWA_B200_L2_FETCH=32" "BWA_B200_KMER_K=11 BWA_B200_L2_FETCH=32"
BENCH_ARGS='--genome 1000000000' bash tools/gpu_try.sh "BWA_B200_KMER_K=0 BWA_B200_L2_FETCH=32" "BWA_B200_KMER_K=11 BWA_B200_L2_FETCH=32" "BWA_B200_KMER_K=13 BWA_B200_L2_FETCH=32"
grep -h "fetch gran" gpurun_out/try.err | head -2
