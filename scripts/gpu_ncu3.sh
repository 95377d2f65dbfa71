#!/bin/bash
# ncu capture of the 3-D fused kernel on the C3-sized problem (one pass per launch)
TAG=${1:-ncu3d}; OUT=gpurun_out/$TAG; mkdir -p $OUT
XINV_FUSED_PPL=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:xm3_std3d -s 6 -c 1 -o $OUT/fused3d_full \
    python scripts/prof_c3.py 12 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
