#!/bin/bash
# ncu launch list of one forward: scripts/r2_launches.sh <workload> <B> <flags>
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 26 -c 13 --csv --log-file gpurun_out/r2_launches_tmp.csv python scripts/quick_time.py $1 $2 $3 > gpurun_out/quick_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_tmp.csv')) if len(r)>5]
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and r[hdr.index('Metric Name')]=='gpu__time_duration.sum':
        print(r[hdr.index('Kernel Name')][:60], r[hdr.index('Metric Value')])
PY
