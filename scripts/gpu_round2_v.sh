#!/bin/bash
mkdir -p gpurun_out
for c in c3o0 "c5 --n-aln 3000" c1; do
python bench.py --config $c --steps 5 --warmup 3 --no-extra --no-cpu-baseline --no-mstep 2> /dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['name'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['kernels_ms'], d['config'].get('mean_sweeps_per_window'), d['parity_checked'])"
done
