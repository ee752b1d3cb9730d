#!/bin/bash
# Kernel time of variant builds of the library on one fixture (run on the GPU box).
# usage: ab_libs.sh <fixture> <shapes> lib1.so lib2.so ...
fx=$1; sh=$2; shift 2
for l in "$@"; do
  echo "== $l"
  QR_B200_LIB=$PWD/$l python tools/tune_shapes.py $fx $sh 2>&1 | tail -n $(echo $sh | tr ',' '\n' | wc -l) | cut -c1-110
done
