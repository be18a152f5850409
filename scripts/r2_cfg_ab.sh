#!/bin/bash
# which earlier block of bench.py changes the configs block's numbers?
for extra in "--no-e2e --no-variants" "--no-e2e" "--no-variants" ""; do
  python bench.py --no-cpu-baseline --no-config4 --no-config5 --steps 20 $extra > gpurun_out/b.json
  echo -n "[$extra] "
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/b.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],3), [(c["workload"][:10], round(c["ms_per_step"],3)) for c in d["configs"]])
PY
done
