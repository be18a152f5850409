#!/bin/bash
# tile pooling vs list pooling on the three rigs (whole step, B = 4 and B = 1)
for wl in MultiviewC MultiviewX Wildtrack; do
  for B in 4 1; do
    echo -n "$wl B=$B tile: "; timeout 120 python scripts/quick_time.py $wl $B 0 2>&1 | tail -1 | cut -c50-130
    echo -n "$wl B=$B list: "; VFA_POOL_TILE=0 timeout 120 python scripts/quick_time.py $wl $B 0 2>&1 | tail -1 | cut -c50-130
  done
done
