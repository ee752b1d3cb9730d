#!/bin/bash
# One ncu --set full capture (with source counters) of the render kernel of a library build.
# usage: ncu_cap.sh <tag> <lib.so> <shape> [fixture]      -> gpurun_out/prof_<tag>.ncu-rep
tag=$1; lib=$2; sh=$3; fx=${4:-demo03_1080p_a4g}
QR_B200_LIB=$PWD/$lib QR_B200_SHAPE=$sh ncu --set full --import-source on --clock-control none \
    -k regex:qr_render -s 2 -c 1 -f -o gpurun_out/prof_$tag python tools/prof_run.py $fx 3 2>&1 | tail -3
