#!/bin/bash
# N GPUs: default bench under torchrun (strong scaling of C3), then the multi-GPU C-ABI tests
N=${1:-8}
mkdir -p gpurun_out
SECONDS=0
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r02_n$N.json 2> gpurun_out/bench_r02_n$N.err; echo "bench n$N exit $? after $SECONDS s"; tail -3 gpurun_out/bench_r02_n$N.err
SECONDS=0
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/pytest_multi_n$N.log 2>&1; echo "pytest exit $? after $SECONDS s"; tail -3 gpurun_out/pytest_multi_n$N.log
