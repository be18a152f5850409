#!/bin/bash
for v in 0; do
echo "== tile_variant $v"
VFA_B200_LIB=$PWD/build/variants/libvfa_prof.so VFA_TILE_VARIANT=$v VFA_UMMA_VARIANT=128 timeout 120 python scripts/quick_time.py MultiviewC 4 0 2>&1 | tail -13
done
