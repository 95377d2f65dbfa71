#!/bin/bash
OUT=gpurun_out/${1:-ncu2}; mkdir -p $OUT
XINV_FUSED_RC_VARIANT=2 ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/rc2_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_rc2.log 2>&1
XINV_FUSED_RC_VARIANT=4 ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/rc4_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_rc4.log 2>&1
XINV_FUSED_RC=0 ncu --set full --clock-control none --import-source on -k regex:xm_std2d -s 4 -c 1 -o $OUT/g2_full \
    python bench.py --steps 1 --warmup 1 --sweeps 20 --cpu-sweeps 2 > $OUT/ncu_g2.log 2>&1
ls -la $OUT
