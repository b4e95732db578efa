#!/bin/bash
# mem_flt_chained_seeds on the device: the long-read tests, then the whole GPU suite
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py -x -q -m gpu > gpurun_out/pytest_long.log 2>&1; echo "align rc=$?"; tail -15 gpurun_out/pytest_long.log
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
