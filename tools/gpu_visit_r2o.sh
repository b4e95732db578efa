#!/bin/bash
# the default bench exactly as the driver launches it at N = 2 (every leg: c2, reseed, cigar, c3 with the index built on the box, c4, c5, CPU arms)
set -u
mkdir -p gpurun_out
t0=$(date +%s)
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2_full.json 2>gpurun_out/bench_n2_full.err; echo "bench n2 rc=$? $(( $(date +%s) - t0 )) s"
tail -4 gpurun_out/bench_n2_full.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2_full.json').read().strip().splitlines()[-1])
sm=d['sub_metrics']
print('N=2 value %.1f M/s e2e %.1f M/s' % (d['value']/1e6, d['e2e']['value']/1e6))
print('c3', sm['c3'].get('reads_per_s'), sm['c3'].get('e2e_reads_per_s'), sm['c3'].get('unavailable'))
print('c4', sm['c4_extension_sweep']['min_GCUPS_per_gpu'], sm['c4_extension_sweep']['max_GCUPS_per_gpu'])
print('c5', sm['c5_seeding']['modes'])
print('reseed', sm['chained_reseed']['reads_per_s'], sm['chained_reseed']['e2e_reads_per_s'])
PY
