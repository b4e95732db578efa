#!/bin/bash
# new ext_pair_kernel (8.5 ALU instructions per column pair, band-sized ring state): parity, measured pipe rates, A/B of the merge form and the ring
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_align.py -m gpu -q -x > gpurun_out/pytest_ext.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_ext.log
tail -4 gpurun_out/pytest_ext.log
python -c "
import json, __graft_entry__ as ge
pkg = ge.load_package()
print(json.dumps(pkg.measure_int_alu(0)))" > gpurun_out/int_alu_r2g.json 2>gpurun_out/int_alu_r2g.err; cat gpurun_out/int_alu_r2g.json
for cfg in "A=1" "BWA_B200_PAIR_MERGE=prmt" "BWA_B200_PAIR_NO_RING=1"; do
  env $cfg timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-c3 --no-cpu-baseline 2>gpurun_out/try.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
c=d['sub_metrics']['chained']; f=d['sub_metrics']['fused_one_seed']
print('$cfg: chained %.2f ms ext %.3f ms %.0f GCUPS frac %.3f | fused %.2f ms ext %.3f' % (c['ms_per_step'], c['kernel_ms']['ext_phase'], c['extension_GCUPS'], d['roofline_extension']['frac'], f['ms_per_step'], f['kernel_ms']['ext_phase']))
print('   bins', {k[16:]: round(v,3) for k,v in f['kernel_ms_bins_serialised'].items() if k.startswith('ext_pair')})
" || tail -5 gpurun_out/try.err
done 2>&1 | tee gpurun_out/try_pair.txt
for cfg in "A=1" "BWA_B200_PAIR_NO_RING=1"; do
  echo "== $cfg"
  env $cfg timeout 900 python tools/sweep_c4_c5.py --only-c4 --jobs 1048576 --reps 3 2>gpurun_out/sweep.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
for r in d['c4_extension']: print('q %5d w %3d  %7.1f GCUPS  %6.2f Mjobs/s' % (r['qlen'], r['w'], r['GCUPS'], r['Mjobs_per_s']))
" || tail -5 gpurun_out/sweep.err
done 2>&1 | tee gpurun_out/sweep_pair.txt
