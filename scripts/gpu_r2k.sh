#!/bin/bash
OUT=gpurun_out/${1:-r2k}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_cluster.py -q -x --timeout 60 2>&1 | tail -4
python scripts/prof_c1.py 2000 | tail -1
python scripts/prof_c1.py 2000 180 360 extend | tail -1
for k in 4 6 12; do XINV_CLUSTER_K=$k timeout 60 python scripts/prof_c1.py 2000 90 180 | tail -1; done
for r in 2 4 8 16; do XINV_CLUSTER_R=$r timeout 60 python scripts/prof_c1.py 2000 46 72 | tail -1; done
