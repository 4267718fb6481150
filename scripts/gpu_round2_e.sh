#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_jit.py tests/test_gpu_memo.py tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > gpurun_out/pytest_r02e.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_r02e.log
python bench.py --config c3o0 --steps 5 --no-extra > gpurun_out/bench_r02_c3o0_jit.json 2> gpurun_out/bench_r02_c3o0_jit.err; echo "bench c3o0 exit $?"
python bench.py --config c5 --n-aln 3000 --steps 3 --no-extra --no-mstep > gpurun_out/bench_r02_c5_jit.json 2> gpurun_out/bench_r02_c5_jit.err; echo "bench c5 exit $?"
python bench.py --config c1 --steps 20 --no-extra > gpurun_out/bench_r02_c1.json 2> gpurun_out/bench_r02_c1.err; echo "bench c1 exit $?"
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mstep --no-extra --no-parity --config c3o0"
$B > gpurun_out/plain_r02_c3o0j.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02_c3o0j.csv $B > gpurun_out/ncu_launch_r02_c3o0j.log 2>&1
echo "launch list exit $?"
$B > gpurun_out/bench_r02_c3o0j.json 2>/dev/null &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_o0j -s 3 -c 1 -f -o gpurun_out/prof_r02_c3o0j $B > gpurun_out/ncu_full_r02_c3o0j.log 2>&1
echo "full capture exit $?"
