#!/bin/bash
# run-time compiled order-0 kernel: parity tests, then kernel time against the precompiled F81 kernel on c3o0 and c5 shapes
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jit.py -x -q > gpurun_out/pytest_r02c.log 2>&1; echo "pytest exit $?"; tail -15 gpurun_out/pytest_r02c.log
for K in 1 2; do
  python scripts/run_o0_only.py c3o0 2000 $K > gpurun_out/jit_c3o0_$K.log 2>&1; tail -2 gpurun_out/jit_c3o0_$K.log | cut -c1-220
  python scripts/run_o0_only.py c5 600 $K > gpurun_out/jit_c5_$K.log 2>&1; tail -2 gpurun_out/jit_c5_$K.log | cut -c1-220
done
