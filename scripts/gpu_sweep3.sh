#!/bin/bash
# dense C3: every 3-D kernel variant x level split; dependent-issue latencies (scripts/micro/lat.cu)
OUT=gpurun_out/${1:-sweep3}; mkdir -p $OUT
nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/lat scripts/micro/lat.cu && /tmp/lat > $OUT/lat.txt 2>&1; cat $OUT/lat.txt
SWEEP_NTZ=1,2,3 python scripts/sweep_c3.py 200 > $OUT/sweep_dense.txt 2>&1; cat $OUT/sweep_dense.txt
