#!/bin/bash
# shape flags from cut_jobs_kernel; host-to-host step over workers x chunks; ncu --set full of the chaining-stage kernels
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
for cfg in "2 2" "2 4" "3 3" "3 6" "4 4" "4 8"; do
  set -- $cfg
  timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline --e2e-workers $1 --e2e-chunks $2 > gpurun_out/bench_e2e_w$1_c$2.json 2>gpurun_out/bench_e2e_w$1_c$2.err; echo "bench w$1 c$2 rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_e2e_w$1_c$2.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']
print('workers $1 chunks $2', 'value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'ext_phase %.3f cut %.3f' % (c['kernel_ms']['ext_phase'], c['kernel_ms']['cut_jobs_kernel']))
PY
done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'chain_kernel|jobs_kernel|cut_jobs_kernel|key_kernel|finish_kernel|fill_kernel|locate_kernel' -s 28 -c 7 \
   -o gpurun_out/prof_r02_chain_stage -f python bench.py --steps 1 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/prof_chain_stage.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r02_chain_stage.ncu-rep
