#!/bin/bash
# final state of round 2: GPU suite, the default bench as the driver launches it at N = 1, the reference arm, the launch list of one C2 step,
# memcheck over the extension tests (closed-form path with two substitutions)
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
t0=$(date +%s)
timeout 2400 python bench.py --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_n1_full.json 2>gpurun_out/bench_n1_full.err; echo "bench n1 rc=$? $(( $(date +%s) - t0 )) s"
tail -3 gpurun_out/bench_n1_full.err
t0=$(date +%s)
timeout 1200 python bench.py --impl reference --gpus 1 --steps 20 --warmup 3 > gpurun_out/bench_ref_full.json 2>gpurun_out/bench_ref_full.err; echo "bench ref rc=$? $(( $(date +%s) - t0 )) s"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final.csv python bench.py --steps 2 --warmup 1 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_extension_matches_oracle or pipeline_matches" > gpurun_out/memcheck_ext.log 2>&1; echo "memcheck ext rc=$?"; tail -3 gpurun_out/memcheck_ext.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n1_full.json').read().strip().splitlines()[-1])
sm=d['sub_metrics']
print('N=1 value %.2f M/s e2e %.2f M/s (one at a time %.2f) ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['one_batch_at_a_time']/1e6, d['ms_per_step']))
print('roofline', {k: d['roofline'][k] for k in ('kernel','achieved','peak','frac')}, 'ext', d['roofline_extension'].get('frac_s16x2'), d['roofline_extension'].get('achieved_gcups'))
print('c3', sm['c3'].get('reads_per_s'), sm['c3'].get('e2e_reads_per_s'), sm['c3'].get('closed_form_jobs'))
print('c4', sm['c4_extension_sweep']['min_GCUPS_per_gpu'], sm['c4_extension_sweep']['max_GCUPS_per_gpu'])
print('c5', sm['c5_seeding']['modes'])
print('reseed', sm['chained_reseed']['reads_per_s'], sm['chained_reseed']['e2e_reads_per_s'])
print('cigar', sm['cigar']['jobs_per_s'], sm['cigar']['e2e_jobs_per_s'])
print('sw', sm['mate_rescue_sw']['jobs_per_s_kernel'], sm['mate_rescue_sw']['e2e_jobs_per_s'])
print('cpu', d['cpu_baseline'])
r=json.loads(open('gpurun_out/bench_ref_full.json').read().strip().splitlines()[-1]); print('ref', r['value'])
PY
