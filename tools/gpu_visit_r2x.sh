#!/bin/bash
# (1) L2 sector lookups per random 32-byte load (ncu over the random-sector benchmark); (2) seeding-kernel occupancy variants in the
# HBM regime (config 3: 3.1 Gb index), where C2's choices (fwd 8 / back 10 blocks per SM) were never re-measured
set -u
mkdir -p gpurun_out
bash tools/l2_sector_count.sh > gpurun_out/l2_sector_count.txt 2>&1; cat gpurun_out/l2_sector_count.txt
for cfg in "8 10" "8 12" "8 8" "10 12" "12 12"; do
  set -- $cfg
  BWA_B200_FWD_MINB=$1 BWA_B200_BACK_MINB=$2 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_c3_f$1_b$2.json 2>gpurun_out/bench_c3_f$1_b$2.err; echo "bench fwd $1 back $2 rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_f$1_b$2.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']['kernel_ms']; c3=d['sub_metrics']['c3']
print('fwd $1 back $2  C2: step %.3f fwd %.3f back %.3f | C3: step %.3f fwd %.3f back %.3f locate %.3f  %.2f M/s' % (d['ms_per_step'], c['fwd_kernel'], c['back_kernel'], c3['ms_per_step'], c3['kernel_ms']['fwd_kernel'], c3['kernel_ms']['back_kernel'], c3['kernel_ms']['locate_kernel'], c3['reads_per_s']/1e6))
PY
done
