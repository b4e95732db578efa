#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
for cfg in "8 10" "8 12" "8 8"; do
  set -- $cfg
  BWA_B200_FWD_MINB=$1 BWA_B200_BACK_MINB=$2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['sub_metrics']['kernel_ms']
print('fwd_minb $1 back_minb $2: step %.2f ms  fwd %.3f back %.3f locate %.3f fill %.3f ext %.3f' % (d['ms_per_step'], k['fwd_kernel'], k['back_kernel'], k['locate_kernel'], k['fill_kernel'], sum(v for n,v in k.items() if n.startswith('ext_inter'))))
"
done 2>&1 | tee gpurun_out/tune_seed.txt
