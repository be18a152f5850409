#!/bin/bash
# single frames: quads' texel lists (default at B = 1) vs the staged-tile pooling (VFA_POOL_TILE=2)
for wl in MultiviewC MultiviewX Wildtrack; do
  for t in 1 2; do
    echo -n "VFA_POOL_TILE=$t "; VFA_POOL_TILE=$t python scripts/quick_time.py $wl 1 0 | cut -c7-120
  done
done
