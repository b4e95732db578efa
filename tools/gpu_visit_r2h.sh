#!/bin/bash
# ext_pair_kernel v2 (17 instructions per column pair in the unrolled loop, ring / chunk logic compiled out where not needed): parity, A/B of 4 vs 8 pairs per trip, C4 grid, ncu
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_align.py -m gpu -q -x > gpurun_out/pytest_ext.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_ext.log
tail -4 gpurun_out/pytest_ext.log
for cfg in "A=1" "BWA_B200_PAIR_UNROLL=8"; do
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-c3 --no-cpu-baseline 2>gpurun_out/try.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; f=d['sub_metrics']['fused_one_seed']
print('$cfg: chained %.2f ms ext %.3f ms %.0f GCUPS frac %.3f | fused %.2f ms ext %.3f' % (c['ms_per_step'], c['kernel_ms']['ext_phase'], c['extension_GCUPS'], d['roofline_extension']['frac'], f['ms_per_step'], f['kernel_ms']['ext_phase']))
print('   bins', {k[16:]: round(v,3) for k,v in f['kernel_ms_bins_serialised'].items() if k.startswith('ext_pair')})
" || tail -5 gpurun_out/try.err
done 2>&1 | tee gpurun_out/try_pair.txt
for cfg in "A=1" "BWA_B200_PAIR_UNROLL=8"; do
  echo "== $cfg"
  env $cfg timeout 900 python tools/sweep_c4_c5.py --only-c4 --jobs 1048576 --reps 3 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['c4_extension']: print('q %5d w %3d  %7.1f GCUPS  %6.2f Mjobs/s' % (r['qlen'], r['w'], r['GCUPS'], r['Mjobs_per_s']))
" || tail -5 gpurun_out/sweep.err
done 2>&1 | tee gpurun_out/sweep_pair.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'ext_pair_kernel' -s 192 -c 16 \
   -o gpurun_out/prof_r02_pair_v2 -f python bench.py --steps 1 --warmup 3 --no-extras --no-c3 --no-cpu-baseline > gpurun_out/prof_pair_v2.log 2>&1; echo "ncu rc=$?"
