#!/bin/bash
for k in 6 12; do XINV_CLUSTER_K=$k python scripts/prof_c1.py 2000 | tail -1; done
for k in 6 12; do XINV_CLUSTER_K=$k python scripts/prof_c1.py 2000 180 360 extend | tail -1; done
XINV_CLUSTER_K=6 timeout 300 python -m pytest tests/test_gpu_cluster.py -q -x --timeout 60 -k "c1_size or land_mask or 6-" 2>&1 | tail -2
