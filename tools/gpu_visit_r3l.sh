#!/bin/bash
# closed form for up to six substitutions (general proof): the GPU suite, then the default bench exactly as the driver launches it
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
t0=$(date +%s)
timeout 1500 python bench.py > gpurun_out/bench_final.json 2>gpurun_out/bench_final.err; echo "bench default rc=$? $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
sm=d['sub_metrics']; c=sm['chained']; c3=sm['c3']; rs=sm['chained_reseed']; lr=sm['long_reads']
print('C2 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'closed', c.get('closed_form_jobs'), 'of', c['jobs_short']+c['jobs_long'], 'cells', c['cells_per_step'], 'gcups', round(c['extension_GCUPS'],1))
print('kernel_ms', {k: round(v,3) for k,v in c['kernel_ms'].items()})
print('C3 %.2f M/s e2e %.2f ms %.3f closed %s ext %.3f' % (c3['reads_per_s']/1e6, c3['e2e_reads_per_s']/1e6, c3['ms_per_step'], c3.get('closed_form_jobs'), c3['kernel_ms']['ext_phase']), 'identical', d['cpu_baseline'].get('gpu_output_identical_on_sample'), c3.get('cpu_baseline',{}).get('gpu_output_identical_on_sample'))
print('reseed %.2f' % (rs['reads_per_s']/1e6), 'long', lr['reads_per_s'], lr['cpu_baseline']['gpu_output_identical_on_sample'])
print('roofline', d['roofline']['frac'], d['roofline_extension']['frac'], 'cpu', d['cpu_baseline']['value'])
PY
