#!/bin/bash
# full bench (both arms) as the driver runs it, then config-3-only variants of the k-mer table depth on the cached 3.1 Gb index
set -u
mkdir -p gpurun_out
( time timeout 1700 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2> gpurun_out/bench_full.time; echo "bench rc=$?"
tail -3 gpurun_out/bench_full.err; cat gpurun_out/bench_full.time
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'e2e', d['e2e']['value'], 'ms', d['ms_per_step'])
    print('roofline', json.dumps(d['roofline'])[:600])
    print('ext', {k: d['roofline_extension'][k] for k in ('achieved_gcups', 'peak_gcups_s16x2', 'frac_s16x2')})
    print('cpu', d['cpu_baseline'])
    print('c3', json.dumps(d['sub_metrics']['c3'])[:1500])
    print('bwa_mem', json.dumps(d['sub_metrics']['bwa_mem_cpu'])[:1200])
except Exception as ex:
    print('parse failed', ex)
PY
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "ref rc=$?"; cat gpurun_out/bench_ref.json | head -c 1500; cat gpurun_out/bench_ref.time
for K in 0 11 12; do
BWA_B200_KMER_K=$K timeout 900 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline --reads 200000 > gpurun_out/bench_c3_k$K.json 2> gpurun_out/bench_c3_k$K.err; echo "c3 K=$K rc=$?"
python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/bench_c3_k$K.json').read().strip().splitlines()[-1])
    c = d['sub_metrics']['c3']
    print('K=$K c3 reads/s', c['reads_per_s'], 'e2e', c['e2e_reads_per_s'], 'ms', c['ms_per_step'], {k: round(v, 3) for k, v in c['kernel_ms'].items()})
except Exception as ex:
    print('parse failed', ex)
PY
done
