#!/bin/bash
# Measurements of the variant builds (multiposenet_b200/build.py: VARIANTS) against the product library.
# usage: bash tools/variants.sh <tag>
T=${1:-r02a}
O=gpurun_out
for v in "" ${VARIANTS:-waves2}; do
  lib=""; name=${v:-product}
  [ -n "$v" ] && lib=$PWD/multiposenet_b200/libmpn_b200_$v.so
  { echo "=== $name"; MPN_LIB=$lib python tools/fused_trace.py 78 2>&1 | tail -15; MPN_LIB=$lib python tools/two_streams.py c2 1 3 2>&1 | tail -2;
    MPN_LIB=$lib python tools/prn_sweep.py 16 78 128 200 256 2>&1 | grep bf16; } > $O/${T}_variant_$name.txt 2>&1
done
cat $O/${T}_variant_*.txt
