#!/bin/bash
# Build a variant of libvfa_b200.so with extra -D switches for A/B timing on the GPU box:
#   scripts/build_variant.sh rows3 -DVFA_QUAD_ROWS=3 -DVFA_QUAD_MINBLOCKS=2   ->  build/variants/libvfa_rows3.so
# Use it with VFA_B200_LIB=build/variants/libvfa_rows3.so python scripts/check_fside.py MultiviewC
set -e
cd "$(dirname "$0")/.."
name=$1; shift
out=build/variants/$name
mkdir -p $out
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in vfa_b200/csrc/*.cu; do
  b=$(basename $f .cu)
  extra=""; [ "$b" = vfa_table ] && extra="-fmad=false"
  if [ "$b" = "${TU:-vfa_fwd_fside}" ]; then
    nvcc -O3 -std=c++17 -lineinfo $ARCH -Xcompiler -fPIC -Xptxas -v "$@" -c $f -o $out/$b.o 2> $out/$b.ptxas.log
  else
    cp build/$b.o $out/$b.o        # unchanged translation units come from the default build
  fi
done
nvcc $ARCH -shared -o build/variants/libvfa_$name.so $out/*.o -cudart static -ldl
grep -E "registers|spill" $out/${TU:-vfa_fwd_fside}.ptxas.log | sort | uniq -c | sort -rn | head -4
