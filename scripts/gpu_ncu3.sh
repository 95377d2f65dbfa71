#!/bin/bash
OUT=gpurun_out/${1:-ncu3}; mkdir -p $OUT
ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/rc2_c2 \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_rc2_c2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/rc2_c5 \
    python bench.py --workload c5 --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_rc2_c5.log 2>&1
ls -la $OUT
