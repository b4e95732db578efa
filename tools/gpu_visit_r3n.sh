#!/bin/bash
# key_kernel's shifted-diagonal test eight rows per step: the GPU suite, C2 + C3 with the CPU identity checks, the launch list of one C2 step
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-c4 --no-c5 > gpurun_out/bench_kw.json 2>gpurun_out/bench_kw.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_kw.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_kw.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; c3=d['sub_metrics']['c3']; rs=d['sub_metrics']['chained_reseed']; lr=d['sub_metrics']['long_reads']
print('C2 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'closed', c.get('closed_form_jobs'), 'of', c['jobs_short']+c['jobs_long'], 'cells', c['cells_per_step'], 'gcups', round(c['extension_GCUPS'],1), 'ext %.3f' % c['kernel_ms']['ext_phase'])
print('C3 %.2f M/s e2e %.2f ms %.3f closed %s ext %.3f' % (c3['reads_per_s']/1e6, c3['e2e_reads_per_s']/1e6, c3['ms_per_step'], c3.get('closed_form_jobs'), c3['kernel_ms']['ext_phase']), 'identical', d['cpu_baseline'].get('gpu_output_identical_on_sample'), c3.get('cpu_baseline',{}).get('gpu_output_identical_on_sample'))
print('reseed %.2f' % (rs['reads_per_s']/1e6), 'long', lr['reads_per_s'], lr['cpu_baseline']['gpu_output_identical_on_sample'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final_v6.csv python bench.py --steps 2 --warmup 1 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/launches_r02_final_v6.csv | grep -i "key_kernel\|back_kernel\|ext_pair_kernel<0, 64, 1, 0, 0"
