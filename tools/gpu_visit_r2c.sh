#!/bin/bash
# k-mer table A/B after a kernel change: seeding parity, then fused-step kernel times on the 100 Mb and 1 Gb indexes
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reseed.py -m gpu -q -x -k "seeding or reseed or pipeline" > gpurun_out/pytest_kt.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_kt.log
tail -4 gpurun_out/pytest_kt.log
rm -f gpurun_out/try.txt
bash tools/gpu_try.sh ${C2_CFGS:-"BWA_B200_KMER_K=0" "BWA_B200_KMER_K=11"}
BENCH_ARGS='--genome 1000000000' bash tools/gpu_try.sh ${G1_CFGS:-"BWA_B200_KMER_K=0" "BWA_B200_KMER_K=11" "BWA_B200_KMER_K=13"}
cp gpurun_out/try.txt gpurun_out/try_kt.txt
