#!/bin/bash
# per-phase cycles of tile_build_kernel (build with: TU=vfa_pool_tile bash scripts/build_variant.sh bprof -DVFA_BUILD_PROFILE)
VFA_B200_LIB=$PWD/build/variants/libvfa_bprof.so timeout 120 python scripts/quick_time.py ${1:-MultiviewC} 4 0 2>&1 | tail -12
