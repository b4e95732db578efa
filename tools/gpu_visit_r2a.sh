#!/bin/bash
# Round 2, first GPU visit: every GPU parity test (incl. the SAM-level one), the composed-stages check, the SAM comparison on
# BASELINE config 1 (reference driver over this library / CPU checker / the reference's own GPU libraries / stock bwa mem),
# the measured roofline denominators, and the human-sized (3.1 Gb) step as the baseline for the seeding work.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/gpu.txt; free -g >> gpurun_out/gpu.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python tools/compose_check.py 3000 > gpurun_out/compose_check.json 2> gpurun_out/compose_check.err; echo "compose rc=$?"; cat gpurun_out/compose_check.json
mkdir -p /tmp/samc1
timeout 900 python tools/sam_check.py --out /tmp/samc1 --threads 4 > gpurun_out/sam_c1.json 2> gpurun_out/sam_c1.err; echo "sam rc=$?"
head -c 1500 gpurun_out/sam_c1.json; for f in /tmp/samc1/*.log; do echo "== $f"; tail -4 $f; done > gpurun_out/sam_c1_logs.txt 2>&1
python -c "
import json, importlib
pkg = importlib.import_module('bwa-mem_gpu_b200')
print(json.dumps(pkg.measure_int_alu(0)))" > gpurun_out/int_alu.json 2>&1; cat gpurun_out/int_alu.json
timeout 300 python tools/rs_sweep.py > gpurun_out/rs_sweep.txt 2>&1; cat gpurun_out/rs_sweep.txt
timeout 1500 python bench.py --genome 3100000000 --reads 1000000 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_base.json 2> gpurun_out/bench_c3_base.err; echo "bench c3 rc=$?"
tail -c 3000 gpurun_out/bench_c3_base.json; tail -5 gpurun_out/bench_c3_base.err
ls -la gpurun_out/
