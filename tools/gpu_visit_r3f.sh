#!/bin/bash
# the long-read leg of the bench (mem_flt_chained_seeds on the device at scale, identity against the reference on a sample)
set -u
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 10 --warmup 3 --no-c3 --no-c4 --no-c5 > gpurun_out/bench_long.json 2>gpurun_out/bench_long.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_long.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_long.json').read().strip().splitlines()[-1])
l=d['sub_metrics']['long_reads']
print({k: l[k] for k in ('reads_per_s','e2e_reads_per_s','ms_per_step','regions_per_step','jobs_short','jobs_long','closed_form_jobs','cells_per_step','extension_GCUPS')})
print({k: round(v,3) for k,v in l['kernel_ms'].items()}); print(l.get('cpu_baseline'))
print('C2 %.2f e2e %.2f' % (d['value']/1e6, d['e2e']['value']/1e6))
PY
