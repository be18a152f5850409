#!/bin/bash
# what the driver runs at round end: GPU tests, smoke, the default bench line
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -6
timeout 600 python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_r2_final.err
python - <<'PY'
import json
d=json.loads([l for l in open("gpurun_out/bench_r2_final.json") if l.startswith("{")][-1])
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], {k:v["value"] for k,v in d["e2e"]["variants"].items()})
print({k:(v.get("value") or v) for k,v in d["variants"].items()}, [c["value"] for c in d["configs"]], d["config4"]["value"], d["config5"].get("value") or d["config5"])
r=d["roofline"]; print(r["kernel"], r["frac"], r["step"]["frac"], r["step"]["frac_tf32x3"], [(k["kernel"],round(k["kernel_ms"],3)) for k in r["kernels"]])
PY
