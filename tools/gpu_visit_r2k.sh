#!/bin/bash
# the default bench, timed end to end
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python bench.py > gpurun_out/bench.json 2>gpurun_out/bench.err; echo "bench rc=$? $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value %.1f M/s  e2e %.1f M/s  frac_ext %.3f  roofline %s' % (d['value']/1e6, d['e2e']['value']/1e6, d['roofline_extension']['frac'], d['roofline']['frac']))
c3=d['sub_metrics']['c3']; print('c3', c3.get('reads_per_s'), c3.get('e2e_reads_per_s'), c3.get('kernel_ms'))
c4=d['sub_metrics']['c4_extension_sweep']; print('c4 min frac', c4['min_frac_s16x2'], 'min/max GCUPS', c4['min_GCUPS_per_gpu'], c4['max_GCUPS_per_gpu'], c4.get('cpu_baseline'))
print('c5', d['sub_metrics']['c5_seeding'])
print('cigar', {k: v for k, v in d['sub_metrics']['cigar'].items() if 'jobs_per_s' in k})
PY
timeout 600 python tools/sa_intv_ab.py > gpurun_out/sa_intv_ab.json 2>gpurun_out/sa_intv_ab.err; tail -4 gpurun_out/sa_intv_ab.err
