#!/bin/bash
# compute-sanitizer memcheck over the kernels added this round (sw_stripe_kernel, seedsw_kernel / chain_long_kernel, key_kernel's closed
# form), then the per-bin times of the extension launch set after the closed-form jobs left it
set -u
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_sw.py -q -m gpu -x -k "materescue or tiny or asym or registers" > gpurun_out/memcheck_sw.log 2>&1; echo "memcheck sw rc=$?"; tail -4 gpurun_out/memcheck_sw.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_align.py -q -m gpu -x -k "long_reads" > gpurun_out/memcheck_long.log 2>&1; echo "memcheck long rc=$?"; tail -4 gpurun_out/memcheck_long.log
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "test_extension_matches_oracle" > gpurun_out/memcheck_ext.log 2>&1; echo "memcheck ext rc=$?"; tail -4 gpurun_out/memcheck_ext.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_sw.py -q -m gpu -x -k "materescue or tiny" > gpurun_out/racecheck_sw.log 2>&1; echo "racecheck sw rc=$?"; tail -4 gpurun_out/racecheck_sw.log
BWA_B200_BENCH_PROFILE_MODE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-c3 --no-c4 --no-c5 --no-cpu-baseline > gpurun_out/bench_bins.json 2>gpurun_out/bench_bins.err; echo "bench bins rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_bins.json').read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']['kernel_ms']
print('step', d['ms_per_step']); print({k: round(v,3) for k,v in c.items()})
PY
