#!/bin/bash
# quick iteration: tile parity test + ncu launch list of one B = 4 MultiviewC forward (and optionally Wildtrack)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_frame_parity.py -x -q -m gpu 2>&1 | tail -4
for wl in MultiviewC ${EXTRA_WL}; do
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 28 -c 14 --csv --log-file gpurun_out/r2_launches_quick_$wl.csv python scripts/quick_time.py $wl 4 0 > gpurun_out/quick_ncu.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_launches_quick_$wl.csv')) if len(r)>5]
hdr=None
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and r[hdr.index('Metric Name')]=='gpu__time_duration.sum':
        print('$wl', r[hdr.index('Kernel Name')][:50], r[hdr.index('Metric Value')])
PY
done
timeout 120 python scripts/quick_time.py MultiviewC 4 0
