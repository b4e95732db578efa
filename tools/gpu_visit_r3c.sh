#!/bin/bash
# two batches in flight in the dispatcher: the compact / multi tests, then the bench's host-to-host step both ways
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_compact.py -q -m gpu -x > gpurun_out/pytest_compact.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_compact.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_stream.json 2>gpurun_out/bench_stream.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_stream.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_stream.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; c3=d['sub_metrics']['c3']
print('C2 value %.2f M/s  e2e streamed %.2f  one at a time %.2f  full records %.2f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['one_batch_at_a_time']/1e6, d['e2e']['full_records_one_batch_in_flight']/1e6))
print('C3 value %.2f M/s  e2e streamed %.2f  one at a time %.2f' % (c3['reads_per_s']/1e6, c3['e2e_reads_per_s']/1e6, c3['e2e_one_batch_at_a_time_reads_per_s']/1e6))
PY
