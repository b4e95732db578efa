#!/bin/bash
# closed form for up to three substitutions: the GPU suite, then C2 + C3 with the CPU identity checks
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-c4 --no-c5 > gpurun_out/bench_cf3.json 2>gpurun_out/bench_cf3.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_cf3.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_cf3.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; c3=d['sub_metrics']['c3']; rs=d['sub_metrics']['chained_reseed']; lr=d['sub_metrics']['long_reads']
print('C2 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'closed', c.get('closed_form_jobs'), 'of', c['jobs_short']+c['jobs_long'], 'cells', c['cells_per_step'], 'gcups', round(c['extension_GCUPS'],1), 'ext %.3f' % c['kernel_ms']['ext_phase'])
print('C3 %.2f M/s e2e %.2f ms %.3f closed %s ext %.3f' % (c3['reads_per_s']/1e6, c3['e2e_reads_per_s']/1e6, c3['ms_per_step'], c3.get('closed_form_jobs'), c3['kernel_ms']['ext_phase']), 'identical', d['cpu_baseline'].get('gpu_output_identical_on_sample'), c3.get('cpu_baseline',{}).get('gpu_output_identical_on_sample'))
print('reseed %.2f' % (rs['reads_per_s']/1e6), 'long', lr['reads_per_s'], lr['cpu_baseline']['gpu_output_identical_on_sample'])
PY
