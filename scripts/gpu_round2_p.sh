#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_train.py tests/test_gpu_mstep.py tests/test_gpu_tree_learning.py tests/test_gpu_multimotif.py tests/test_gpu_order1.py tests/test_gpu_alignment_models.py tests/test_gpu_xmlio.py -q > gpurun_out/pytest_opt_r02p.log 2>&1; echo "pytest exit $?"; tail -40 gpurun_out/pytest_opt_r02p.log
python scripts/time_mstep.py c3 2>&1 | tail -2 | cut -c1-400
python scripts/time_mstep.py c2 2>&1 | tail -1 | cut -c1-400
