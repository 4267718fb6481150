#!/bin/bash
# tuning: C2 bench for a list of extra bench arguments (one quoted string per run)
mkdir -p gpurun_out
i=0
for extra in "$@"; do
  i=$((i+1))
  timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-mstep $extra > gpurun_out/ab2_$i.json 2> gpurun_out/ab2_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab2_$i.json"))
print("$extra", round(d["ms_per_step"],4), round(d["value"]/1e9,3), "e2e", round(d["e2e"]["value"]/1e9,3), [(k["kernel"][:24], k["ms"]) for k in d["kernels_ms"]])
PY
done
