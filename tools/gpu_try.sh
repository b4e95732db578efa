#!/bin/bash
# BENCH_ARGS='--genome 1000000000' adds bench.py arguments
# usage: gpu_try.sh "ENV=.. ENV=.." ...   -> one short bench per env set, kernel times printed
set -u
mkdir -p gpurun_out
for cfg in "$@"; do
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-chain ${BENCH_ARGS:-} 2>gpurun_out/try.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
k=d['sub_metrics']['kernel_ms']
print('$cfg: step %.2f ms e2e %.1f Mr/s fwd %.3f back %.3f locate %.3f fill %.3f cut %.3f ext %.3f' % (d['ms_per_step'], d['e2e']['value']/1e6, k['fwd_kernel'], k['back_kernel'], k['locate_kernel'], k['fill_kernel'], k['cut_kernel'], sum(v for n,v in k.items() if n.startswith('ext_'))))
" || tail -5 gpurun_out/try.err
  grep -h "L2 persist" gpurun_out/try.err | head -1
done 2>&1 | tee -a gpurun_out/try.txt
