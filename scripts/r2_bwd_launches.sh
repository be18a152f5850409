#!/bin/bash
# ncu launch list of one forward + backward step (B = 4): scripts/r2_bwd_launches.sh [workload]
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 150 -c 80 --csv --log-file gpurun_out/r2_bwd_launches.csv python scripts/time_bwd_kernels.py ${1:-MultiviewC} > gpurun_out/bwd_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_bwd_launches.csv')) if len(r)>5]
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and r[hdr.index('Metric Name')]=='gpu__time_duration.sum':
        print(r[hdr.index('Kernel Name')][:70], r[hdr.index('Metric Value')])
PY
python scripts/time_bwd_kernels.py ${1:-MultiviewC}
