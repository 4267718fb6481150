#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_jit.py tests/test_gpu_target_shapes.py tests/test_gpu_parity.py -q -x > gpurun_out/pytest_jit_r02u.log 2>&1; echo "pytest exit $?"; tail -4 gpurun_out/pytest_jit_r02u.log
for c in c3o0 "c5 --n-aln 3000" c1; do
python bench.py --config $c --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-mstep --no-parity 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['name'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels_ms'])"
done
