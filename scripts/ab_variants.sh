#!/bin/bash
# A/B of compile-time variants of the pooling kernel (scripts/build_variant.sh) on the GPU box:
#   scripts/ab_variants.sh "MultiviewC Wildtrack" default rows2 rows3 ...
wl=$1; shift
for v in "$@"; do
  lib=""; [ "$v" != default ] && lib=$PWD/build/variants/libvfa_$v.so
  echo "=== $v"
  VFA_B200_LIB=$lib timeout 120 python scripts/check_fside.py $wl 2>&1 | grep -v "^$" | cut -c1-260
done
