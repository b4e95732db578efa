#!/bin/bash
# ncu --set full of sw_stripe_kernel on the mate-rescue bench workload
set -u
mkdir -p gpurun_out
cat > /tmp/run_sw.py <<'PY'
import sys, json, importlib, argparse
sys.path.insert(0, '.')
import bench
pkg = importlib.import_module("bwa-mem_gpu_b200")
args = argparse.Namespace(no_cpu_baseline=True, steps=5)
print(json.dumps(bench.run_mate_sw(args, pkg)))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'sw_stripe_kernel' -s 1 -c 1 -o gpurun_out/prof_r02_sw_stripe -f python /tmp/run_sw.py > gpurun_out/prof_sw.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/prof_sw.log
