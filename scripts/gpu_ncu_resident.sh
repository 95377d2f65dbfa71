#!/bin/bash
# ncu capture of the resident kernel on a 73x37 nine-point (invert_Eliassen-sized) section
TAG=${1:-ncur}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xr_resident -s 1 -c 1 -o $OUT/resident_full \
    python scripts/prof_resident.py 300 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
