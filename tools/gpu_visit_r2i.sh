#!/bin/bash
# compact boundary + multi dispatcher parity; A/B of 8 vs 16 pairs per trip; e2e through the dispatcher with 2 / 4 / 8 chunks per batch
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_compact.py tests/test_gpu_parity.py tests/test_gpu_align.py -m gpu -q -x > gpurun_out/pytest_cmp.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_cmp.log
tail -15 gpurun_out/pytest_cmp.log
for cfg in "A=1" "BWA_B200_PAIR_UNROLL=16" "E2E=2" "E2E=8"; do
  extra=""
  case $cfg in E2E=*) extra="--e2e-chunks ${cfg#E2E=}";; esac
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-c3 --no-cpu-baseline $extra 2>gpurun_out/try.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; f=d['sub_metrics']['fused_one_seed']
print('$cfg: chained %.2f ms ext %.3f ms %.0f GCUPS frac %.3f | e2e compact %.1f M/s (%d chunks) full-records 1 batch %.1f M/s | fused %.2f ms ext %.3f' % (c['ms_per_step'], c['kernel_ms']['ext_phase'], c['extension_GCUPS'], d['roofline_extension']['frac'], c['e2e_reads_per_s']/1e6, c['e2e_chunks_per_step'], c['e2e_full_records_one_batch_in_flight']/1e6, f['ms_per_step'], f['kernel_ms']['ext_phase']))
print('   bins', {k[16:]: round(v,3) for k,v in f['kernel_ms_bins_serialised'].items() if k.startswith('ext_pair')})
" || tail -5 gpurun_out/try.err
done 2>&1 | tee gpurun_out/try_pair.txt
for cfg in "BWA_B200_PAIR_UNROLL=16"; do
  echo "== $cfg"
  env $cfg timeout 900 python tools/sweep_c4_c5.py --only-c4 --jobs 1048576 --reps 3 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['c4_extension']: print('q %5d w %3d  %7.1f GCUPS  %6.2f Mjobs/s' % (r['qlen'], r['w'], r['GCUPS'], r['Mjobs_per_s']))
" || tail -5 gpurun_out/sweep.err
done 2>&1 | tee gpurun_out/sweep_pair.txt
