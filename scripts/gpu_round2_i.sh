#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02i.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -5 gpurun_out/pytest_gpu_r02i.log
SECONDS=0
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02i.log 2>&1; echo "smoke exit $? after $SECONDS s"
SECONDS=0
python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err; echo "bench exit $? after $SECONDS s"; tail -3 gpurun_out/bench_r02_default.err
