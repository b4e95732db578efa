#!/bin/bash
# closed-form extension jobs + mem_flt_chained_seeds on the device: the GPU suite, then the C2 chained step with and without the shortcut
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu.log
for mode in on off; do
  if [ $mode = off ]; then export BWA_B200_EXT_NO_CLOSED=1; else unset BWA_B200_EXT_NO_CLOSED; fi
  timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 > gpurun_out/bench_closed_$mode.json 2>gpurun_out/bench_closed_$mode.err; echo "bench $mode rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_closed_$mode.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']
print('$mode', 'value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'closed', c.get('closed_form_jobs'), 'cells', c['cells_per_step'], 'gcups', c['extension_GCUPS'])
print({k: round(v,3) for k,v in c['kernel_ms'].items()})
print('identical', d['cpu_baseline'].get('gpu_output_identical_on_sample'))
PY
done
