#!/bin/bash
# ncu launch list of one forward (B = 4 MultiviewC) + full capture of the pooling kernel and its list builder
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 26 -c 26 --csv --log-file gpurun_out/r2_launches_quick.csv python scripts/quick_time.py MultiviewC 4 0 > gpurun_out/quick_ncu.log 2>&1
grep -o '"[a-z_A-Z0-9<>:, ]*kernel[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","[^"]*","gpu__time_duration.sum","[a-z]*","[0-9.,]*"' gpurun_out/r2_launches_quick.csv | sed 's/"[0-9]*","[^"]*","[^"]*","[^"]*","[^"]*",//' | head -40
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"pool_tile_kernel|tile_build_kernel" -s 4 -c 2 -o gpurun_out/r2_pool_tile -f python scripts/quick_time.py MultiviewC 4 0 > gpurun_out/ncu_pool_tile.log 2>&1
ls -la gpurun_out/r2_pool_tile.ncu-rep
