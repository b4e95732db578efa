#!/bin/bash
# ext_wave_kernel: parity (long / wide-score jobs of banded batches), then the long-read points of the extension sweep with and without it
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "extension" > gpurun_out/pytest_wave.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_wave.log
tail -25 gpurun_out/pytest_wave.log
for cfg in "A=1" "BWA_B200_EXT_NO_WIDE=1" "BWA_B200_EXT_NO_WAVE=1"; do
  echo "== $cfg"
  env $cfg timeout 600 python tools/sweep_c4_c5.py --only-long --reps 3 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['c4_extension']: print('q %5d w %3d  %7.1f GCUPS  %6.2f Mjobs/s  %d jobs' % (r['qlen'], r['w'], r['GCUPS'], r['Mjobs_per_s'], r['jobs']))
" || tail -5 gpurun_out/sweep.err
done 2>&1 | tee gpurun_out/sweep_wave.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
