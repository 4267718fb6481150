#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_mstep.py tests/test_gpu_train.py tests/test_gpu_tree_learning.py tests/test_gpu_multimotif.py tests/test_gpu_multi.py -q -x > gpurun_out/pytest_mstep_r02r.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_mstep_r02r.log
for i in 1 2; do python scripts/time_mstep.py c3 2>&1 | tail -3 | cut -c1-330; done
