#!/bin/bash
# one multi-GPU bench line: scripts/r2_scale.sh N   (under gpurun --gpus N)
N=$1
timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 5 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
tail -2 gpurun_out/bench_r2_n$N.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/bench_r2_n$N.json") if l.startswith("{")][-1])
print("dp", d["value"], d["ms_per_step"])
s=d["strong"]; print("strong", {k:s[k] for k in ("value","ms_per_step","speedup","efficiency","cameras_per_rank","collective")})
print("config4", {k:d["config4"][k] for k in ("value","ms_per_step","frames_per_rank")})
print("config5", {k:d["config5"].get(k) for k in ("value","ms_per_step","error")})
print("e2e", d["e2e"]["value"], {k:v["value"] for k,v in d["e2e"]["variants"].items()})
print("variants", {k:v["value"] for k,v in d["variants"].items()})
PY
