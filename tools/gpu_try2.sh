#!/bin/bash
# GPU parity tests + one bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q ${PYTEST_K:-} > gpurun_out/pytest_try.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_try.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_try.json 2> gpurun_out/bench_try.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_try.err
python -c "
import json
d=json.load(open('gpurun_out/bench_try.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline_extension'])
k=d['sub_metrics']['kernel_ms']
print({a:round(v,3) for a,v in k.items() if v>0.02}, 'sum', sum(k.values()))"
