#!/bin/bash
# AB order >= 1 + multimotif tests; L2 access-policy window on/off: time and DRAM traffic of the order-1 kernel and of the 30-species order-0 kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_alignment_models.py tests/test_gpu_multimotif.py -x -q > gpurun_out/pytest_r02b.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_r02b.log
for P in 0 1; do
  python scripts/run_o1_only.py 2000 2 3 -1 $P > gpurun_out/l2p_o1_$P.log 2>&1; tail -2 gpurun_out/l2p_o1_$P.log | cut -c1-200
  python scripts/run_o0_only.py c5 600 -1 $P > gpurun_out/l2p_c5_$P.log 2>&1; tail -2 gpurun_out/l2p_c5_$P.log | cut -c1-200
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:score_o --csv --log-file gpurun_out/l2p_o1_traffic_$P.csv python scripts/run_o1_only.py 2000 2 1 -1 $P > /dev/null 2>&1
  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:score_o --csv --log-file gpurun_out/l2p_c5_traffic_$P.csv python scripts/run_o0_only.py c5 600 -1 $P > /dev/null 2>&1
done
grep -h "score_o" gpurun_out/l2p_*traffic*.csv | cut -d, -f5,13- | head -40
