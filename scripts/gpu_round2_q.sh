#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_mstep.py -q -x > gpurun_out/pytest_mstep_r02q.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_mstep_r02q.log
for i in 1 2; do python scripts/time_mstep.py c3 2>&1 | tail -3 | cut -c1-330; done
