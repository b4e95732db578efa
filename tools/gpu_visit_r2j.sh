#!/bin/bash
# whole GPU suite, the default bench timed end to end (both arms), launch list + ncu --set full of the shipped ext_pair_kernel bins
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
t0=$(date +%s)
timeout 1500 python bench.py --impl reference > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; echo "ref rc=$? $(( $(date +%s) - t0 )) s"
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
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ext_pair_kernel' -s 192 -c 16 \
   -o gpurun_out/prof_r02_pair_v3 -f python bench.py --steps 1 --warmup 3 --no-extras --no-c3 --no-c4 --no-cpu-baseline > gpurun_out/prof_pair_v3.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 2 --warmup 3 --no-extras --no-c3 --no-c4 --no-cpu-baseline > gpurun_out/launches_r02.log 2>&1; echo "launch list rc=$?"
