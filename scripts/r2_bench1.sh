#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 5 > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo rc=$?; tail -3 gpurun_out/bench_r2_a.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r2_a.json').read().strip().splitlines()[-1])
r=d.pop('roofline'); 
print(json.dumps({k:v for k,v in d.items() if k not in ('e2e',)},indent=None)[:1800])
print('e2e', json.dumps(d['e2e'])[:1200])
print('step', json.dumps(r['step'])[:900])
for k in r['kernels']: print({x:k[x] for x in ('kernel','kernel_ms','achieved','frac','kernel_share_of_step')})
PY
