#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tree_learning.py -q -x -k heterogeneity > gpurun_out/pytest_het_r02z.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_het_r02z.log
