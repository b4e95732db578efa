#!/bin/bash
# two GPUs: the two-device dispatcher test, then the bench at N = 2 (short: C2 legs only) and the single-process dispatcher over both devices
set -u
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_compact.py -m gpu -q > gpurun_out/pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_2gpu.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras --no-c3 --no-c4 --no-cpu-baseline > gpurun_out/bench_n2.json 2>gpurun_out/bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.1f M/s e2e %.1f M/s full-records-1-batch %.1f' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['full_records_one_batch_in_flight']/1e6))
PY
timeout 900 python tools/multi_single_process.py > gpurun_out/multi_single_process.json 2>gpurun_out/multi_single_process.err; tail -3 gpurun_out/multi_single_process.err; cat gpurun_out/multi_single_process.json
