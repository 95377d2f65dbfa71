#!/bin/bash
# ncu capture of the cluster kernel on the C1-sized problem + quick timings over (R, K)
TAG=${1:-ncuc}; OUT=gpurun_out/$TAG; mkdir -p $OUT
python scripts/prof_c1.py 2000 | tail -1
python scripts/prof_c1.py 2000 180 360 extend | tail -1
XINV_CLUSTER_R=8 XINV_CLUSTER_K=16 python scripts/prof_c1.py 2000 90 180 | tail -1
for k in 4 6 8 12; do XINV_CLUSTER_K=$k python scripts/prof_c1.py 2000 90 180 | tail -1; done
for r in 1 2 4 8 16; do XINV_CLUSTER_R=$r python scripts/prof_c1.py 2000 46 72 | tail -1; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:xc_cluster -s 1 -c 1 -o $OUT/cluster_full \
    python scripts/prof_c1.py 300 > $OUT/ncu.log 2>&1; echo "ncu rc=$?"; tail -2 $OUT/ncu.log
