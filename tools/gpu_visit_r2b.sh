#!/bin/bash
# Round 2, second GPU visit: the k-mer interval table -- parity of every seeding consumer, then A/B timings on the 100 Mb (L2 regime)
# and 1 Gb (HBM regime) indexes for several table depths.
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reseed.py tests/test_gpu_align.py tests/test_compat_driver.py -m gpu -q -x > gpurun_out/pytest_kt.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_kt.log
tail -15 gpurun_out/pytest_kt.log
rm -f gpurun_out/try.txt
bash tools/gpu_try.sh "BWA_B200_KMER_K=0" "BWA_B200_KMER_K=10" "BWA_B200_KMER_K=11" "BWA_B200_KMER_K=12"
BENCH_ARGS='--genome 1000000000' bash tools/gpu_try.sh "BWA_B200_KMER_K=0" "BWA_B200_KMER_K=11" "BWA_B200_KMER_K=12" "BWA_B200_KMER_K=13"
cp gpurun_out/try.txt gpurun_out/try_kt.txt
