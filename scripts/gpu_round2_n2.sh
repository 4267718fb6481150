#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench n2 exit $?"; tail -2 gpurun_out/bench_r02_n2.err
