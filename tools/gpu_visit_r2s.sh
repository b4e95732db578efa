#!/bin/bash
# chain_kernel occupancy variants; the seeding kernels on an index that fits one L2 partition (bounds what a denser bucket can give on C2)
set -u
mkdir -p gpurun_out
for mb in 1 8 10 12 16; do
  BWA_B200_CHAIN_MINB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_minb$mb.json 2>gpurun_out/bench_minb$mb.err; echo "bench minb $mb rc=$?"
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_minb$mb.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']
print('minb $mb', 'value %.2f M/s ms %.3f' % (d['value']/1e6, d['ms_per_step']), {k: round(v,3) for k,v in c['kernel_ms'].items()})
PY
done
for g in 66700000 60000000; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline --genome $g > gpurun_out/bench_g$g.json 2>gpurun_out/bench_g$g.err; echo "bench genome $g rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_g$g.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']
print('genome $g', 'value %.2f M/s ms %.3f' % (d['value']/1e6, d['ms_per_step']), {k: round(v,3) for k,v in c['kernel_ms'].items()}, 'seeds', c['seeds'])
PY
done
