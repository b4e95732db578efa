#!/bin/bash
# ncu --set full of fwd_kernel / back_kernel of one fused bench step with and without the k-mer table (100 Mb index, optionally GENOME=...)
set -u
mkdir -p gpurun_out
G=${GENOME:-100000000}
for K in ${KS:-0 11}; do
BWA_B200_KMER_K=$K timeout 900 ncu --set full --clock-control none --import-source on -k regex:'back_kernel|fwd_kernel' -s 6 -c 2 \
   -o gpurun_out/prof_kt_g${G}_k$K -f python bench.py --genome $G --steps 1 --warmup 3 --no-chain --no-cpu-baseline > gpurun_out/prof_kt_$K.log 2>&1; echo "ncu K=$K rc=$?"
done
ls -la gpurun_out/*.ncu-rep
