#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_bg_orders.py tests/test_gpu_order1.py -x -q > gpurun_out/pytest_bgk_r02l.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_bgk_r02l.log
