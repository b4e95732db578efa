#!/bin/bash
# mate-rescue SW bench leg (baseline kernel) + GPU tests touched by it
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_sw.py -q -m gpu -x > gpurun_out/pytest_sw.log 2>&1; echo "pytest sw rc=$?"; tail -3 gpurun_out/pytest_sw.log
timeout 900 python - <<'PY' > gpurun_out/mate_sw.json 2> gpurun_out/mate_sw.err
import sys, json, importlib, argparse
sys.path.insert(0, '.')
import bench
pkg = importlib.import_module("bwa-mem_gpu_b200")
args = argparse.Namespace(no_cpu_baseline=False, steps=5)
print(json.dumps(bench.run_mate_sw(args, pkg)))
PY
echo "mate_sw rc=$?"; tail -2 gpurun_out/mate_sw.err; cat gpurun_out/mate_sw.json
