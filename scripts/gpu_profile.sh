#!/bin/bash
# ncu evidence for the bench command (run only after the same command exited 0 without ncu, in the same call).
# Usage: bash scripts/gpu_profile.sh <tag> [extra bench args]
TAG=${1:-r01}; shift
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline $*"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_o -s 3 -c 2 -f -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_full_$TAG.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -12
