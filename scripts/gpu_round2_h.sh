#!/bin/bash
mkdir -p gpurun_out
SECONDS=0
python bench.py > gpurun_out/bench_r02_default.json 2> gpurun_out/bench_r02_default.err; echo "bench exit $? after $SECONDS s"
SECONDS=0
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02_reference.json 2> gpurun_out/bench_r02_reference.err; echo "reference exit $? after $SECONDS s"
