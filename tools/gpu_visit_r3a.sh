#!/bin/bash
# closed form for up to two substitutions: the GPU suite, then C2 + C3 with and without it
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
for mode in on off; do
  if [ $mode = off ]; then export BWA_B200_EXT_NO_CLOSED=1; else unset BWA_B200_EXT_NO_CLOSED; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-c4 --no-c5 > gpurun_out/bench_cf2_$mode.json 2>gpurun_out/bench_cf2_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_cf2_$mode.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; c3=d['sub_metrics']['c3']
print('$mode', 'C2 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'closed', c.get('closed_form_jobs'), 'of', c['jobs_short']+c['jobs_long'], 'cells', c['cells_per_step'], 'gcups', round(c['extension_GCUPS'],1), 'ext %.3f' % c['kernel_ms']['ext_phase'])
print('   C3 %.2f M/s e2e %.2f ms %.3f closed %s ext %.3f' % (c3['reads_per_s']/1e6, c3['e2e_reads_per_s']/1e6, c3['ms_per_step'], c3.get('closed_form_jobs'), c3['kernel_ms']['ext_phase']), 'identical', d['cpu_baseline'].get('gpu_output_identical_on_sample'), c3.get('cpu_baseline',{}).get('gpu_output_identical_on_sample'))
PY
done
