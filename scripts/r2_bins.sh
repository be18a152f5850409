#!/bin/bash
# tile order: work quantised to N levels (locality inside a level) vs exact order; pool time + DRAM bytes
for lib in bins4 bins8 bins16 bins65536; do
  l=$PWD/build/variants/libvfa_$lib.so
  [ -f $l ] || continue
  VFA_B200_LIB=$l timeout 300 ncu --metrics dram__bytes_read.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:"pool_tile_kernel" -s 3 -c 2 --csv --log-file gpurun_out/bins_$lib.csv python scripts/quick_time.py MultiviewC 4 0 > /dev/null 2>&1
  echo -n "$lib: "; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/bins_$lib.csv')) if len(r)>5]
hdr=None; out=[]
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr: out.append((r[hdr.index('Metric Name')].split('.')[0][-12:], r[hdr.index('Metric Value')]))
print(out)
PY
  VFA_B200_LIB=$l python scripts/quick_time.py MultiviewC 4 0 | cut -c40-100
done
