#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_jit.py -q -x 2>&1 | tail -2
python bench.py --config c5 --n-aln 3000 --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-mstep 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['name'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels_ms'], d['parity_checked']['sweep_counts'])"
