#!/bin/bash
# key_kernel: the shifted diagonals of a closed-form candidate tested by the whole warp.  Extension / pipeline / aligner GPU tests, C2 with the
# CPU identity check, launch list.
set -u
mkdir -p gpurun_out
timeout 130 python -m pytest tests/test_gpu_parity.py tests/test_gpu_align.py -q -m gpu -x -k "extension or pipeline or align" > gpurun_out/pytest_kw.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_kw.log
timeout 100 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 > gpurun_out/bench_kw2.json 2>gpurun_out/bench_kw2.err; echo "bench rc=$?"; tail -1 gpurun_out/bench_kw2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_kw2.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']
print('C2 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), 'closed', c.get('closed_form_jobs'), 'of', c['jobs_short']+c['jobs_long'], 'cells', c['cells_per_step'], 'ext %.3f' % c['kernel_ms']['ext_phase'], 'identical', d['cpu_baseline'].get('gpu_output_identical_on_sample'))
PY
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final_v7.csv python bench.py --steps 2 --warmup 1 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
python tools/launch_summary.py gpurun_out/launches_r02_final_v7.csv | grep -i "key_kernel"
