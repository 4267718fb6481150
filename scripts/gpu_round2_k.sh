#!/bin/bash
# 2 GPUs: the multi-GPU tests (skipped on one GPU) and the default bench under torchrun, then the reference arm
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_r02k.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_multi_r02k.log
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_r02_n2.json 2> gpurun_out/bench_r02_n2.err; echo "bench n2 exit $? after $SECONDS s"; tail -3 gpurun_out/bench_r02_n2.err
