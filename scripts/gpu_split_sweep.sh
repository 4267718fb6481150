#!/bin/bash
# tuning: C2 bench for several values of pattern_memo_split
mkdir -p gpurun_out
for s in 0 8 12 16 20 24 32; do
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --memo-split $s > gpurun_out/split_$s.json 2> gpurun_out/split_$s.err
  python - <<PY
import json
d=json.load(open("gpurun_out/split_$s.json"))
print($s, round(d["ms_per_step"],4), round(d["value"]/1e9,3), [(k["kernel"][:28], k["ms"]) for k in d["kernels_ms"]])
PY
done
