#!/bin/bash
# seedsw_kernel on the register-resident 16-bit stripe pass: long-read parity tests, then the long-read bench leg
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_align.py -q -m gpu -x > gpurun_out/pytest_align.log 2>&1; echo "pytest align rc=$?"; tail -4 gpurun_out/pytest_align.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_align.py -q -m gpu -x -k "long_reads_given" > gpurun_out/memcheck_long.log 2>&1; echo "memcheck long rc=$?"; tail -3 gpurun_out/memcheck_long.log
bash tools/gpu_visit_r3f.sh
