#!/bin/bash
# round-2 evidence: GPU tests, the default bench line, launch list + one full ncu capture of the dominant kernel of C3 and c3o0
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r02a.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_r02a.log
python bench.py > gpurun_out/bench_r02_c3.json 2> gpurun_out/bench_r02_c3.err; echo "bench exit $?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mstep --no-extra --no-parity"
for cfg in c3 c3o0; do
  K=score_o1n; [ $cfg = c3o0 ] && K=score_o0f
  $B --config $cfg > gpurun_out/plain_r02_$cfg.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_$cfg.csv $B --config $cfg > gpurun_out/ncu_launch_r02_$cfg.log 2>&1
  echo "launch list $cfg exit $?"
  $B --config $cfg > gpurun_out/bench_r02p_$cfg.json 2>/dev/null &&
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o gpurun_out/prof_r02_$cfg $B --config $cfg > gpurun_out/ncu_full_r02_$cfg.log 2>&1
  echo "full capture $cfg exit $?"
done
ls -la gpurun_out | tail -12
