#!/bin/bash
# per-kernel times of the list-based pooling variants (bench.py's CUDA-event measurement) + an ncu launch list
for v in default lb2 lb4m2 lb6m2; do
  for l32 in 0 1; do
  lib=""; [ "$v" != default ] && lib=$PWD/build/variants/libvfa_$v.so
  VFA_POOL_LIST32=$l32 VFA_B200_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload ${WL:-MultiviewC} 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v lanes32=$l32', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'pool_ms', round(r['kernel_ms'],3), 'gemm_ms', round(r['second_kernel']['kernel_ms'],3))"
  done
done
VFA_POOL_LIST32=1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_list.csv python scripts/quick_time.py MultiviewC 4 0 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_list.csv')) if len(r)>5 and r[0].isdigit()]
from collections import defaultdict
agg=defaultdict(list)
for r in rows: agg[r[4].split('(')[0][:60]].append(float(r[-1]))
for k,v in agg.items(): print(f'{k:60s} n={len(v)} last={v[-1]:.1f}')
P
