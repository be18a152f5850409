#!/bin/bash
VFA_TILE_VARIANT=${TV:-3} VFA_UMMA_VARIANT=128 timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pool_tile_kernel" -s 2 -c 1 -o gpurun_out/r2_pool_tile_v${TV:-3} -f python scripts/quick_time.py MultiviewC 4 0 > gpurun_out/ncu_pool_tile_v3.log 2>&1
ls -la gpurun_out/r2_pool_tile_v${TV:-3}.ncu-rep
