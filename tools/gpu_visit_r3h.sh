#!/bin/bash
# smoke() as the driver runs it, then the default bench (every leg, the long-read leg included)
set -u
mkdir -p gpurun_out
timeout 900 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
t0=$(date +%s)
timeout 2400 python bench.py > gpurun_out/bench_default.json 2>gpurun_out/bench_default.err; echo "bench default rc=$? $(( $(date +%s) - t0 )) s"
tail -2 gpurun_out/bench_default.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1])
sm=d['sub_metrics']
print('N=1 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']))
print('long', sm['long_reads']['reads_per_s'], sm['long_reads'].get('cpu_baseline'))
print('c3', sm['c3'].get('reads_per_s'), sm['c3'].get('e2e_reads_per_s'))
print(list(sm.keys()))
PY
