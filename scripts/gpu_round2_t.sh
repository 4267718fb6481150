#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bg_orders.py -q -x -k "training" > gpurun_out/pytest_bgtrain_r02t.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_bgtrain_r02t.log
