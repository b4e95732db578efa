#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list, ncu --set full of the seeding
# kernels and of every extension bin of one timed step.  Outputs in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E 'Model name|^CPU\(s\)' >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
STEPS=${STEPS:-10}
timeout 1500 python bench.py --steps $STEPS --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; echo "ncu list rc=$?"
# 5 seeding-side kernels per step (fwd, back, fill, locate, cut): skip the 3 warm-up steps, capture the timed one
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'back_kernel|fwd_kernel|fill_kernel|locate_kernel|cut_kernel' -s 15 -c 5 \
   -o gpurun_out/prof_seed -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_seed.log 2>&1; echo "ncu seed rc=$?"
timeout 1200 ncu --set full --clock-control none --import-source on \
   -k regex:'ext_pair_kernel' -s ${PAIR_SKIP:-36} -c ${PAIR_COUNT:-12} \
   -o gpurun_out/prof_pair -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_pair.log 2>&1; echo "ncu pair rc=$?"
ls -la gpurun_out/
