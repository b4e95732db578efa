#!/bin/bash
# GPU tests of the extension + one bench
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "extension or pipeline" > gpurun_out/pytest_ext.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_ext.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_try.json 2> gpurun_out/bench_try.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_try.json'))
print(d['ms_per_step'], d['value'], d['roofline_extension'])
print({k:round(v,3) for k,v in d['sub_metrics']['kernel_ms'].items() if v>0.02})"
