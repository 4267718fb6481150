#!/bin/bash
# tuning: memo tests, then the C2 bench for a list of "split:wave" settings
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_memo.py -x -q -m gpu 2>&1 | tail -15
for sw in ${@:-16:1 16:0 0:1 0:0}; do
  s=${sw%%:*}; w=${sw##*:}
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --memo-split $s --memo-wave $w > gpurun_out/ab_${s}_$w.json 2> gpurun_out/ab_${s}_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${s}_$w.json"))
print("split $s wave $w", round(d["ms_per_step"],4), round(d["value"]/1e9,3), "e2e", round(d["e2e"]["value"]/1e9,3), [(k["kernel"][:24], k["ms"]) for k in d["kernels_ms"]])
PY
done
