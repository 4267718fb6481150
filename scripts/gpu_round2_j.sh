#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tree_learning.py tests/test_gpu_train.py tests/test_gpu_mstep.py -x -q > gpurun_out/pytest_tree_r02j.log 2>&1; echo "pytest exit $?"; tail -30 gpurun_out/pytest_tree_r02j.log
timeout 600 python scripts/time_mstep.py c3 > gpurun_out/time_mstep_c3.log 2>&1; tail -4 gpurun_out/time_mstep_c3.log
