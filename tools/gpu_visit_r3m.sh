#!/bin/bash
# final build: smoke() as the driver runs it, the new closed-form GPU test (plain and under memcheck), the launch list of one C2 step,
# and the host-to-host step with other worker / chunk counts
set -u
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "closed_form_flanks" > gpurun_out/pytest_cf.log 2>&1; echo "pytest cf rc=$?"; tail -2 gpurun_out/pytest_cf.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "closed_form_flanks and kw0" > gpurun_out/memcheck_cf.log 2>&1; echo "memcheck cf rc=$?"; tail -3 gpurun_out/memcheck_cf.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02_final_v5.csv python bench.py --steps 2 --warmup 1 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/launches_bench.log 2>&1; echo "ncu launches rc=$?"
for wc in "2 2" "3 3" "2 4" "3 6"; do
  set -- $wc
  timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline --e2e-workers $1 --e2e-chunks $2 > gpurun_out/bench_e2e_w$1_c$2.json 2>/dev/null
  python - "$1" "$2" <<'PY'
import json, sys
d=json.loads(open('gpurun_out/bench_e2e_w%s_c%s.json' % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print('workers', sys.argv[1], 'chunks', sys.argv[2], 'value %.2f e2e %.2f one-at-a-time %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['one_batch_at_a_time']/1e6))
PY
done
