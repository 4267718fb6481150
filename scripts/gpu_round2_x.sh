#!/bin/bash
bash scripts/gpu_profile.sh r02f_c3 score_o1n --no-extra --no-mstep --no-parity
cp gpurun_out/plain_r02f_c3.log gpurun_out/bench_r02f_c3.json
