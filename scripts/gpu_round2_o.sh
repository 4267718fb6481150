#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 2400 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/pytest_gpu_r02o.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -45 gpurun_out/pytest_gpu_r02o.log
