#!/bin/bash
# per-kernel times of the list-based pooling (bench.py's CUDA-event measurement) for the default build and the variant
# libraries under build/variants/ (scripts/build_variant.sh), + an ncu launch list of one forward
for v in default ${VARIANTS}; do
  lib=""; [ "$v" != default ] && lib=$PWD/build/variants/libvfa_$v.so
  for wl in ${WL:-MultiviewC}; do
  VFA_B200_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --workload $wl 2>&1 | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']
print('$v $wl', 'value', round(d['value'],1), 'ms/step', round(d['ms_per_step'],3), 'pool_ms', round(r['kernel_ms'],3), 'gemm_ms', round(r['second_kernel']['kernel_ms'],3))"
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_list.csv python scripts/quick_time.py MultiviewC 4 0 > /dev/null 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_list.csv')) if len(r)>5 and r[0].isdigit()]
from collections import defaultdict
agg=defaultdict(list)
for r in rows: agg[r[4].split('(')[0][:60]].append(float(r[-1]))
for k,v in agg.items(): print(f'{k:60s} n={len(v)} last={v[-1]:.1f}')
P
