#!/bin/bash
# Multi-GPU bench check: bash scripts/gpu_multi.sh <N> <tag>   (run under `gpurun --gpus N`)
N=${1:-2}; TAG=${2:-r01}
mkdir -p gpurun_out
P=$((29500 + RANDOM % 1000))
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
echo "bench N=$N exit $?"; cat gpurun_out/bench_${TAG}_n$N.json; tail -5 gpurun_out/bench_${TAG}_n$N.err
timeout 300 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_n1.json 2> gpurun_out/bench_${TAG}_n1.err
echo "bench N=1 exit $?"; cat gpurun_out/bench_${TAG}_n1.json
