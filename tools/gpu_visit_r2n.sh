#!/bin/bash
# eight GPUs: the bench at N = 8 (C2 legs only) and the single-process dispatcher over all devices
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l; nproc; nvidia-smi topo -m | head -12
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-cpu-baseline > gpurun_out/bench_n8.json 2>gpurun_out/bench_n8.err; echo "bench n8 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n8.json').read().strip().splitlines()[-1])
print('N=8 value %.1f M/s e2e %.1f M/s full-records-1-batch %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['full_records_one_batch_in_flight']/1e6))
PY
timeout 900 python tools/multi_single_process.py > gpurun_out/multi_single_process_8.json 2>gpurun_out/multi_single_process_8.err; tail -3 gpurun_out/multi_single_process_8.err
