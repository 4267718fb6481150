#!/bin/bash
# One gpurun call: GPU parity tests, smoke, a short bench, then the ncu launch list of the same bench command.
# Usage (from the repo root, on the GPU box): bash scripts/gpu_check.sh [tag]
TAG=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$TAG.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu_$TAG.log
tail -15 gpurun_out/pytest_gpu_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1
echo "smoke exit $?"; tail -3 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
echo "bench exit $?"; cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
