#!/bin/bash
# What the driver runs at round end, on one fresh box: the GPU tests, smoke(), the default bench and the reference arm.
mkdir -p gpurun_out
SECONDS=0
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_r02_final.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -4 gpurun_out/pytest_gpu_r02_final.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_r02_final.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/smoke_r02_final.log
SECONDS=0
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; echo "bench exit $? after $SECONDS s"; tail -3 gpurun_out/bench_r02_final.err
SECONDS=0
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_r02_final_reference.json 2> gpurun_out/bench_r02_final_reference.err; echo "reference exit $? after $SECONDS s"
