#!/bin/bash
# cut_jobs_kernel: 32 / 16 / 8 lanes per job
set -u
mkdir -p gpurun_out
for lpj in 32 16 8; do
  BWA_B200_CUT_LPJ=$lpj timeout 600 python -m pytest tests/test_gpu_align.py -q -m gpu -x > gpurun_out/pytest_lpj$lpj.log 2>&1; echo "pytest lpj $lpj rc=$?"; tail -1 gpurun_out/pytest_lpj$lpj.log
  BWA_B200_CUT_LPJ=$lpj timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_lpj$lpj.json 2>gpurun_out/bench_lpj$lpj.err; echo "bench lpj $lpj rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_lpj$lpj.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']['kernel_ms']
print('lpj $lpj  step %.3f cut %.3f' % (d['ms_per_step'], c['cut_jobs_kernel']))
PY
done
