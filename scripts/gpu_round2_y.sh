#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scan.py tests/test_golden.py -m gpu -q -x > gpurun_out/pytest_scan_r02y.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_scan_r02y.log
