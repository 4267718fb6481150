#!/bin/bash
# ncu --set full of the trajectory kernels, wavefront on and off (after the plain run of the same command exited 0)
mkdir -p gpurun_out
for w in 1 0; do
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --memo-wave $w"
  $CMD > gpurun_out/plain_w$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:memo_traj -s 6 -c 2 -f -o gpurun_out/prof_traj_w$w $CMD > gpurun_out/ncu_traj_w$w.log 2>&1
  echo "wave $w capture exit $?"
done
ls -la gpurun_out/*.ncu-rep
