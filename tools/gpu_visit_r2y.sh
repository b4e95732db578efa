#!/bin/bash
# chain_kernel with shared-memory scratch: parity; then back_kernel / fwd_kernel at 6 / 7 / 8 blocks per SM on C2 and C3
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py tests/test_gpu_reseed.py tests/test_gpu_compact.py -q -m gpu -x > gpurun_out/pytest_align.log 2>&1; echo "pytest align rc=$?"; tail -4 gpurun_out/pytest_align.log
for cfg in "8 8" "8 7" "8 6" "6 8"; do
  set -- $cfg
  BWA_B200_FWD_MINB=$1 BWA_B200_BACK_MINB=$2 timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_c3_f$1_b$2.json 2>gpurun_out/bench_c3_f$1_b$2.err; echo "bench fwd $1 back $2 rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_c3_f$1_b$2.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']['kernel_ms']; c3=d['sub_metrics']['c3']
print('fwd $1 back $2  C2: step %.3f fwd %.3f back %.3f chain %.3f | C3: step %.3f fwd %.3f back %.3f chain %.3f  %.2f M/s' % (d['ms_per_step'], c['fwd_kernel'], c['back_kernel'], c['chain_kernel'], c3['ms_per_step'], c3['kernel_ms']['fwd_kernel'], c3['kernel_ms']['back_kernel'], c3['kernel_ms']['chain_kernel'], c3['reads_per_s']/1e6))
PY
done
