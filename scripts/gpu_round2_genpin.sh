#!/bin/bash
# The device side of the generator pins and the opt-in HKY (one short call; 8 tests, ~4 s on the box).
mkdir -p gpurun_out
timeout 110 python -m pytest tests/test_generator_pins.py tests/test_gpu_parity.py -m gpu -k "bundled_backgrounds or hky" -q > gpurun_out/pytest_genpin.log 2>&1
echo "exit $?"; tail -5 gpurun_out/pytest_genpin.log
