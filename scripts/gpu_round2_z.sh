#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -q -x -k bundled > gpurun_out/pytest_em_r02z.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_em_r02z.log
