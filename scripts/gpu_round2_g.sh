#!/bin/bash
mkdir -p gpurun_out
# order-1 kernel: warps per SM 8 / 12 / 16 (larger launch bounds), L2 window on the q rows only
for Q in 8 12 16; do
  PHYFOO_VERBOSE=1 python scripts/run_o1_only.py 2000 2 3 $Q > gpurun_out/o1n_q$Q.log 2>&1; grep -h "phyfoo: L2" gpurun_out/o1n_q$Q.log | head -1; tail -2 gpurun_out/o1n_q$Q.log | cut -c1-200
done
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:score_o1n --csv --log-file gpurun_out/o1n_traffic_q8.csv python scripts/run_o1_only.py 2000 2 1 8 > /dev/null 2>&1
grep -h "score_o" gpurun_out/o1n_traffic_q8.csv | cut -d, -f13-
# compiled order-0 kernel: block size caps
for T in 384 512 576 672 768; do
  python scripts/run_o0_only.py c3o0 2000 2 1 $T > gpurun_out/jit_T$T.log 2>&1; echo "T=$T"; tail -2 gpurun_out/jit_T$T.log | cut -c1-120
done
for T in 288 384 512; do
  python scripts/run_o0_only.py c5 600 2 1 $T > gpurun_out/jit5_T$T.log 2>&1; echo "c5 T=$T"; tail -2 gpurun_out/jit5_T$T.log | cut -c1-120
done
