#!/bin/bash
# 8 GPUs: BASELINE configs[3] at full size (3 motifs + background, 100 000 alignments x 1 kb, 12 500 per GPU)
mkdir -p gpurun_out
SECONDS=0
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 8 --config c4 --steps 5 --warmup 3 > gpurun_out/bench_r02_c4_n8.json 2> gpurun_out/bench_r02_c4_n8.err; echo "bench c4 n8 exit $? after $SECONDS s"; tail -3 gpurun_out/bench_r02_c4_n8.err; cut -c1-1200 gpurun_out/bench_r02_c4_n8.json
