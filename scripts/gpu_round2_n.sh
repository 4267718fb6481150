#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order1.py tests/test_gpu_target_shapes.py tests/test_gpu_train.py -x -q > gpurun_out/pytest_o1_r02n.log 2>&1; echo "pytest exit $?"; tail -5 gpurun_out/pytest_o1_r02n.log
python bench.py --config c3 --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-mstep > gpurun_out/bench_r02n_c3.json 2> gpurun_out/bench_r02n_c3.err; echo "bench exit $?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02n_c3.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernels_ms'], d['parity_checked'])"
