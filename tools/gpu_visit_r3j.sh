#!/bin/bash
# back_kernel candidate-change gate: parity at 0 / 6, then C2 + C3 timings at 0 / 4 / 8 / 12
set -u
mkdir -p gpurun_out
for g in 0 6; do
BWA_B200_BACK_GATE=$g timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reseed.py -q -m gpu -x -k "seed or reseed or pipeline" > gpurun_out/pytest_gate$g.log 2>&1; echo "pytest gate $g rc=$?"; tail -1 gpurun_out/pytest_gate$g.log
done
for g in 0 4 8 12; do
  BWA_B200_BACK_GATE=$g timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_gate$g.json 2>gpurun_out/bench_gate$g.err; echo "bench gate $g rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_gate$g.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']['kernel_ms']; c3=d['sub_metrics']['c3']
print('gate $g  C2: step %.3f back %.3f | C3: step %.3f back %.3f' % (d['ms_per_step'], c['back_kernel'], c3['ms_per_step'], c3['kernel_ms']['back_kernel']))
PY
done
