#!/bin/bash
# branch-free counter / L2 picks in the seeding kernels: parity, then C2 + C3
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reseed.py tests/test_gpu_align.py -q -m gpu -x > gpurun_out/pytest_seed.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_seed.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-extras --no-c4 --no-cpu-baseline > gpurun_out/bench_sel4.json 2>gpurun_out/bench_sel4.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_sel4.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; c3=d['sub_metrics']['c3']; c5=d['sub_metrics']['c5_seeding']
print('C2 value %.2f M/s e2e %.2f M/s ms %.3f' % (d['value']/1e6, d['e2e']['value']/1e6, d['ms_per_step']), {k: round(v,3) for k,v in c['kernel_ms'].items()})
print('C3 %.2f M/s e2e %.2f ms %.3f' % (c3['reads_per_s']/1e6, c3['e2e_reads_per_s']/1e6, c3['ms_per_step']), {k: round(v,3) for k,v in c3['kernel_ms'].items()})
print('C5', c5['modes'])
PY
