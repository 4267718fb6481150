#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_order1.py tests/test_gpu_train.py tests/test_gpu_mstep.py tests/test_gpu_multi.py -q -x > gpurun_out/pytest_sink.log 2>&1; echo "pytest exit $?"; tail -12 gpurun_out/pytest_sink.log
python scripts/time_mstep.py c3 2>&1 | tail -2 | cut -c1-420
python bench.py --config c3 --steps 3 --warmup 3 --no-extra --no-cpu-baseline --no-mstep --no-parity 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['kernels_ms'])"
