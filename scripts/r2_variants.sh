#!/bin/bash
# pool_tile_kernel builds A/B: time of the step without the GEMM (VFA_UMMA_VARIANT=128), full kernel and no-work variant
timeout 600 python -m pytest tests/test_gpu_frame_parity.py -x -q -m gpu 2>&1 | tail -2
for lib in default ${LIBS}; do
  l=""; [ "$lib" != default ] && l=$PWD/build/variants/libvfa_$lib.so
  for v in ${TVS:-0 3}; do
    echo -n "$lib tile_variant=$v: "
    VFA_B200_LIB=$l VFA_TILE_VARIANT=$v VFA_UMMA_VARIANT=128 timeout 120 python scripts/quick_time.py ${WL:-MultiviewC} 4 0 2>&1 | tail -1 | cut -c50-120
  done
done
