#!/bin/bash
# c4 (3-motif mixture) on one GPU at reduced size; fresh ncu capture of C3's order-1 kernel (L2 window on)
mkdir -p gpurun_out
python bench.py --config c4 --n-aln 4000 --steps 3 > gpurun_out/bench_r02_c4_n4000.json 2> gpurun_out/bench_r02_c4_n4000.err; echo "bench c4 exit $?"; tail -3 gpurun_out/bench_r02_c4_n4000.err
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-mstep --no-extra --no-parity --config c3"
$B > gpurun_out/bench_r02b_c3.json 2>/dev/null &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:score_o1n -s 3 -c 1 -f -o gpurun_out/prof_r02b_c3 $B > gpurun_out/ncu_full_r02b_c3.log 2>&1
echo "full capture exit $?"
